"""On-disk volume format of the GT-TSDF path: `full_tsdf_layer{l}.npz`.

The reference writes each fused volume with `np.savez_compressed(path, tsdf_vol)` (`tools/data_gen/scannet.py:115`)
and reads it back with `np.load(path, allow_pickle=True).f.arr_0` (`deep3dmap/datasets/scannet.py:103-105`).  Once
integration runs on the B200 the single-threaded zlib pass over 537 MB (512^3 fp32) is what the wall clock shows, so
this module writes the SAME container -- a zip archive with one deflated member `arr_0.npy` holding a version-1.0
.npy stream -- but compresses it chunk-parallel on the host cores:

  * the raw .npy byte stream is cut into chunks; every chunk is deflated independently (raw deflate, no dictionary
    carried over) and terminated with Z_SYNC_FLUSH, the last one with Z_FINISH.  The concatenation is one valid
    deflate stream (the pigz construction), so `np.load` / `zipfile` / the reference reader see an ordinary .npz;
  * the zip CRC-32 is folded from per-chunk CRCs with the GF(2) combine (no second serial pass);
  * the chunk table travels in a private zip extra field (header id 0x3344) of the member, which lets
    `load_npz` inflate the chunks in parallel again; files written by numpy itself are read through `np.load`.

Arrays read back are bit-identical to what `np.savez_compressed` would have stored; the file bytes are not (deflate
block boundaries differ), which no consumer of the format depends on.
"""
import io
import os
import struct
import time
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_EXTRA_ID = 0x3344          # private extra-field id carrying the chunk table
_ZIP64_ID = 0x0001
_U32_MAX = 0xFFFFFFFF
_MIN_CHUNK = 1 << 20
_MAX_CHUNKS = 8192          # keeps the extra field (20 + 4 n bytes) inside the 64 KiB zip limit


# ---- CRC-32 combine (zlib's crc32_combine restated: advance crc1 through len2 zero bytes in GF(2)) -----------
def _gf2_times(mat, vec):
    s, i = 0, 0
    while vec:
        if vec & 1:
            s ^= mat[i]
        vec >>= 1
        i += 1
    return s


def _gf2_square(mat):
    return [_gf2_times(mat, mat[n]) for n in range(32)]


def crc32_combine(crc1, crc2, len2):
    """CRC-32 of A+B from crc32(A), crc32(B) and len(B)."""
    if len2 <= 0:
        return crc1
    odd = [0xEDB88320] + [1 << n for n in range(31)]   # operator for one zero bit
    even = _gf2_square(odd)                              # two zero bits
    odd = _gf2_square(even)                              # four zero bits
    while True:
        even = _gf2_square(odd)
        if len2 & 1:
            crc1 = _gf2_times(even, crc1)
        len2 >>= 1
        if not len2:
            break
        odd = _gf2_square(even)
        if len2 & 1:
            crc1 = _gf2_times(odd, crc1)
        len2 >>= 1
        if not len2:
            break
    return crc1 ^ crc2


# ---- writer ------------------------------------------------------------------------------------------------------
def _npy_header(arr):
    fp = io.BytesIO()
    np.lib.format.write_array_header_1_0(fp, np.lib.format.header_data_from_array_1_0(arr))
    return fp.getvalue()


def _deflate_chunk(view, level, last):
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    out = co.compress(view) + co.flush(zlib.Z_FINISH if last else zlib.Z_SYNC_FLUSH)
    return out, zlib.crc32(view)


def _dos_time(t=None):
    tm = time.localtime(t)
    return ((tm.tm_hour << 11) | (tm.tm_min << 5) | (tm.tm_sec // 2),
            (max(tm.tm_year, 1980) - 1980) << 9 | (tm.tm_mon << 5) | tm.tm_mday)


def savez_compressed(file, arr, threads=None, level=6, chunk_bytes=None, force_zip64=False):
    """Parallel equivalent of `np.savez_compressed(file, arr)` (one positional array -> member `arr_0`).

    `file` gets the `.npz` suffix appended when missing, exactly like numpy does.  Returns a dict with the byte
    counts and the chunk geometry (used by the bench)."""
    arr = np.asanyarray(arr)
    if arr.dtype.hasobject:
        raise ValueError("object arrays are not part of this format")
    path = os.fspath(file)
    if not path.endswith(".npz"):
        path += ".npz"
    data = np.ascontiguousarray(arr)
    if arr.flags.f_contiguous and not arr.flags.c_contiguous:
        data = np.asfortranarray(arr)       # numpy stores Fortran arrays transposed; header says fortran_order
        body = memoryview(data.T.reshape(-1).view(np.uint8)) if data.size else memoryview(b"")
    else:
        body = memoryview(data.reshape(-1).view(np.uint8)) if data.size else memoryview(b"")
    header = _npy_header(arr)
    usize = len(header) + body.nbytes
    if chunk_bytes is None:
        chunk_bytes = max(_MIN_CHUNK, -(-usize // _MAX_CHUNKS))
    chunk_bytes = int(chunk_bytes)
    # chunk 0 starts with the .npy header so that chunk k covers raw bytes [k*chunk, (k+1)*chunk)
    first = bytes(header) + bytes(body[:max(0, chunk_bytes - len(header))])
    pieces = [memoryview(first)]
    off = max(0, chunk_bytes - len(header))
    if len(header) > chunk_bytes:           # pathological tiny chunk: keep the header whole in chunk 0
        off = 0
    while off < body.nbytes:
        pieces.append(body[off:off + chunk_bytes])
        off += chunk_bytes
    n = len(pieces)
    if n > _MAX_CHUNKS + 1:
        raise ValueError("chunk_bytes too small for %d bytes" % usize)
    threads = threads or min(32, os.cpu_count() or 1)
    if n == 1 or threads == 1:
        done = [_deflate_chunk(p, level, i == n - 1) for i, p in enumerate(pieces)]
    else:
        with ThreadPoolExecutor(threads) as pool:
            done = list(pool.map(lambda ip: _deflate_chunk(ip[1], level, ip[0] == n - 1), enumerate(pieces)))
    crc = 0
    for (_, c), p in zip(done, pieces):
        crc = crc32_combine(crc, c, p.nbytes)
    csize = sum(len(b) for b, _ in done)

    name = b"arr_0.npy"
    zip64 = force_zip64 or usize >= _U32_MAX or csize >= _U32_MAX
    table = struct.pack("<QQI", len(first), chunk_bytes, n) + b"".join(struct.pack("<I", len(b)) for b, _ in done)
    private = struct.pack("<HH", _EXTRA_ID, len(table)) + table
    dtime, ddate = _dos_time()
    local_extra = struct.pack("<HHQQ", _ZIP64_ID, 16, usize, csize) if zip64 else b""
    local = struct.pack("<IHHHHHIIIHH", 0x04034B50, 45 if zip64 else 20, 0, 8, dtime, ddate, crc,
                        _U32_MAX if zip64 else csize, _U32_MAX if zip64 else usize, len(name), len(local_extra))
    central_extra = (struct.pack("<HHQQQ", _ZIP64_ID, 24, usize, csize, 0) if zip64 else b"") + private
    central = struct.pack("<IHHHHHHIIIHHHHHII", 0x02014B50, 45 if zip64 else 20, 45 if zip64 else 20, 0, 8, dtime,
                          ddate, crc, _U32_MAX if zip64 else csize, _U32_MAX if zip64 else usize, len(name),
                          len(central_extra), 0, 0, 0, 0o600 << 16, _U32_MAX if zip64 else 0)
    cd_off = len(local) + len(name) + len(local_extra) + csize
    cd_len = len(central) + len(name) + len(central_extra)
    with open(path, "wb") as f:
        f.write(local)
        f.write(name)
        f.write(local_extra)
        for b, _ in done:
            f.write(b)
        f.write(central)
        f.write(name)
        f.write(central_extra)
        if zip64:
            f.write(struct.pack("<IQHHIIQQQQ", 0x06064B50, 44, 45, 45, 0, 0, 1, 1, cd_len, cd_off))
            f.write(struct.pack("<IIQI", 0x07064B50, 0, cd_off + cd_len, 1))
            f.write(struct.pack("<IHHHHIIH", 0x06054B50, 0, 0, 0xFFFF, 0xFFFF, _U32_MAX, _U32_MAX, 0))
        else:
            f.write(struct.pack("<IHHHHIIH", 0x06054B50, 0, 0, 1, 1, cd_len, cd_off, 0))
    return {"path": path, "raw_bytes": usize, "compressed_bytes": csize, "chunks": n, "chunk_bytes": chunk_bytes,
            "threads": threads, "zip64": bool(zip64)}


# ---- reader ------------------------------------------------------------------------------------------------------
def _chunk_table(extra):
    pos = 0
    while pos + 4 <= len(extra):
        hid, ln = struct.unpack_from("<HH", extra, pos)
        if hid == _EXTRA_ID:
            first_raw, step, n = struct.unpack_from("<QQI", extra, pos + 4)
            return first_raw, step, list(struct.unpack_from("<%dI" % n, extra, pos + 24))
        pos += 4 + ln
    return None


def load_npz(file, threads=None, verify_crc=True):
    """-> the `arr_0` array of a `full_tsdf_layer*.npz` (what `np.load(file, allow_pickle=True).f.arr_0` returns in
    `deep3dmap/datasets/scannet.py:103-105`).  Archives written by `savez_compressed` above are inflated
    chunk-parallel; anything else (e.g. written by numpy) goes through `np.load`."""
    import zipfile
    path = os.fspath(file)
    with zipfile.ZipFile(path) as zf:
        try:
            info = zf.getinfo("arr_0.npy")
        except KeyError:
            info = None
        table = _chunk_table(info.extra) if info is not None and info.compress_type == zipfile.ZIP_DEFLATED else None
    if table is None:
        with np.load(path, allow_pickle=True) as z:
            return z.f.arr_0
    first_raw, step, csizes = table
    with open(path, "rb") as f:
        f.seek(info.header_offset)
        lh = f.read(30)
        fnlen, exlen = struct.unpack_from("<HH", lh, 26)
        f.seek(info.header_offset + 30 + fnlen + exlen)
        blob = f.read(info.compress_size)
    if sum(csizes) != len(blob):
        raise ValueError("%s: chunk table does not match the member size" % path)
    view = memoryview(blob)
    offs = np.concatenate([[0], np.cumsum(csizes)]).astype(np.int64)

    def inflate(i):
        d = zlib.decompressobj(-15)
        return d.decompress(view[offs[i]:offs[i + 1]])

    head = inflate(0)
    if len(head) != first_raw:
        raise ValueError("%s: first chunk inflates to %d bytes, table says %d" % (path, len(head), first_raw))
    fp = io.BytesIO(head)
    version = np.lib.format.read_magic(fp)
    if version != (1, 0):
        raise ValueError("unexpected .npy version %r" % (version,))
    shape, fortran, dtype = np.lib.format.read_array_header_1_0(fp)
    hlen = fp.tell()
    count = int(np.prod(shape, dtype=np.int64))
    out = np.empty(count, dtype=dtype)
    raw = out.view(np.uint8)
    if raw.nbytes + hlen != info.file_size:
        raise ValueError("%s: size mismatch" % path)
    raw[:first_raw - hlen] = np.frombuffer(head, dtype=np.uint8, offset=hlen)
    n = len(csizes)
    crcs = [0] * n
    lens = [0] * n
    crcs[0], lens[0] = zlib.crc32(head), len(head)
    if n > 1:
        # chunk k >= 1 covers raw bytes [first_raw + (k-1)*step, first_raw + k*step)
        def work(i):
            b = inflate(i)
            o = first_raw - hlen + (i - 1) * step
            raw[o:o + len(b)] = np.frombuffer(b, dtype=np.uint8)
            crcs[i], lens[i] = (zlib.crc32(b) if verify_crc else 0), len(b)

        threads = threads or min(32, os.cpu_count() or 1)
        with ThreadPoolExecutor(threads) as pool:
            list(pool.map(work, range(1, n)))
        if sum(lens) != info.file_size:
            raise ValueError("%s: inflated %d bytes, expected %d" % (path, sum(lens), info.file_size))
    if verify_crc:
        crc = 0
        for c, ln in zip(crcs, lens):
            crc = crc32_combine(crc, c, ln)
        if crc != info.CRC:
            raise ValueError("%s: CRC mismatch" % path)
    return out.reshape(shape, order="F" if fortran else "C")
