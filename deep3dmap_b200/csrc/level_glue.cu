// SURVEY §8 f2 -- the steps either side of back_project inside NeuConNet's coarse-to-fine loop
// (models/neucon_network.py:52-89, 113-122, 143-154, 180-207; core/voxel/generate_grids.py:4-11):
// coordinate generation, 1->8 upsampling of coords/features, the world->aligned-camera transform of the sparse
// coordinates, GT look-up, and the occupancy threshold + ORDERED compaction that defines the next level's sparsity.
// All of it is HBM-bound row/index movement: 128-bit rows where the layout allows, coalesced 4-byte words
// otherwise, grids sized from the element count, no float atomics, results independent of scheduling.
#include "d3m_common.cuh"

namespace d3m {

static inline unsigned blocks_for(int64_t n, int per_block) { return (unsigned)((n + per_block - 1) / per_block); }

// ------------------------------------------------------------------------------------------------------------------
// generate_grid (generate_grids.py:4-11) + the per-fragment [b | xyz] rows of neucon_network.py:118-122
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) grid_coords_kernel(int gy, int gz, int64_t n_per, int interval, int B,
                                                          float4* __restrict__ coords, float* __restrict__ grid3) {
  pdl_enter();
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_per * (coords ? B : 1)) return;
  const int b = (int)(t / n_per);
  const int64_t n = t - (int64_t)b * n_per;
  const int z = (int)(n % gz);
  const int64_t r = n / gz;
  const int y = (int)(r % gy);
  const int x = (int)(r / gy);
  const float fx = (float)(x * interval), fy = (float)(y * interval), fz = (float)(z * interval);
  if (coords) coords[t] = make_float4((float)b, fx, fy, fz);
  if (grid3 && b == 0) {
    grid3[n] = fx;
    grid3[n_per + n] = fy;
    grid3[2 * n_per + n] = fz;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// upsample (neucon_network.py:68-89): child i of a voxel adds `interval` to the axes of
// pos_list = [x],[y],[z],[x,y],[x,z],[y,z],[x,y,z] (child 0 = the voxel itself); dtype preserved.
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void child_offsets(int i, int& dx, int& dy, int& dz) {
  // i:      0  1  2  3  4  5  6  7
  // x bit:  0  1  0  0  1  1  0  1   -> 0xB2 ; y: 0 0 1 0 1 0 1 1 -> 0xD4 ; z: 0 0 0 1 0 1 1 1 -> 0xE8
  dx = (0xB2 >> i) & 1;
  dy = (0xD4 >> i) & 1;
  dz = (0xE8 >> i) & 1;
}

template <int KIND>
__global__ void __launch_bounds__(256) upsample_coords_kernel(const void* __restrict__ pre, int64_t total, int num,
                                                              int interval, void* __restrict__ up) {
  pdl_enter();
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int64_t n = t / num;
  const int i = (int)(t - n * num);
  int dx, dy, dz;
  child_offsets(i, dx, dy, dz);
  if (KIND == D3M_COORDS_F32) {
    float4 c = __ldg(reinterpret_cast<const float4*>(pre) + n);
    const float s = (float)interval;
    if (dx) c.y += s;
    if (dy) c.z += s;
    if (dz) c.w += s;
    reinterpret_cast<float4*>(up)[t] = c;
  } else if (KIND == D3M_COORDS_I64) {
    longlong2 a = __ldg(reinterpret_cast<const longlong2*>(pre) + 2 * n);
    longlong2 c = __ldg(reinterpret_cast<const longlong2*>(pre) + 2 * n + 1);
    a.y += dx * interval;
    c.x += dy * interval;
    c.y += dz * interval;
    reinterpret_cast<longlong2*>(up)[2 * t] = a;
    reinterpret_cast<longlong2*>(up)[2 * t + 1] = c;
  } else {
    int4 c = __ldg(reinterpret_cast<const int4*>(pre) + n);
    c.y += dx * interval;
    c.z += dy * interval;
    c.w += dz * interval;
    reinterpret_cast<int4*>(up)[t] = c;
  }
}

// up_feat = pre_feat.unsqueeze(1).expand(-1,num,-1): every input row is written `num` times, back to back.
// One thread = one VecT (128 / 64 / 32 bit, the widest that divides the row and the pointers' alignment) of one INPUT
// row: loaded once, stored to the `num` children rows -- consecutive lanes cover consecutive vectors of a row, so every
// store instruction of a warp writes whole 128-byte lines of one or two output rows.  IdxT = uint32 whenever the vector
// count fits (64-bit division costs ~100 instructions per thread).
template <typename VecT, typename IdxT>
__global__ void __launch_bounds__(256) upsample_feat_kernel(const VecT* __restrict__ pre, IdxT total_vec, IdxT vec_per_row,
                                                            int num, VecT* __restrict__ up) {
  pdl_enter();
  const IdxT t = (IdxT)blockIdx.x * (IdxT)blockDim.x + threadIdx.x;
  if (t >= total_vec) return;
  const IdxT n = t / vec_per_row;
  const IdxT j = t - n * vec_per_row;
  const VecT v = __ldg(pre + t);
  VecT* o = up + ((int64_t)n * num) * vec_per_row + j;
#pragma unroll 8
  for (int i = 0; i < num; ++i) __stcs(o + (int64_t)i * vec_per_row, v);
}

template <typename VecT>
static void launch_upsample_feat(const float* pre, int64_t N, int C, int num, float* up, cudaStream_t stream) {
  const int per = (int)(sizeof(VecT) / 4);
  const int64_t vpr = C / per, total = N * vpr;
  if (total < (1ll << 32)) {
    launch_k(upsample_feat_kernel<VecT, uint32_t>, dim3(blocks_for(total, 256)), dim3(256), 0, stream, 
        reinterpret_cast<const VecT*>(pre), (uint32_t)total, (uint32_t)vpr, num, reinterpret_cast<VecT*>(up));
  } else {
    launch_k(upsample_feat_kernel<VecT, uint64_t>, dim3(blocks_for(total, 256)), dim3(256), 0, stream, 
        reinterpret_cast<const VecT*>(pre), (uint64_t)total, (uint64_t)vpr, num, reinterpret_cast<VecT*>(up));
  }
}

// ------------------------------------------------------------------------------------------------------------------
// world -> aligned-camera coordinates of the sparse voxels (neucon_network.py:143-154):
//   r = [xyz*voxel_size + origin[b], 1] @ W[b,:3,:]^T   (fp32; K=4 dot product as the k=0..3 FMA chain of sgemm)
//   output rows [rx, ry, rz, b]  (":154  r_coords[:, [1,2,3,0]]"); rows of no fragment keep float(coords).
// ------------------------------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256) aligned_camera_kernel(const void* __restrict__ coords, int64_t N,
                                                             const float* __restrict__ origin, int B, float vs,
                                                             const float* __restrict__ w2ac,
                                                             float4* __restrict__ out) {
  pdl_enter();
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float x, y, z, bf;
  const int b = load_coord<KIND>(coords, n, B, x, y, z);
  if (KIND == D3M_COORDS_F32) bf = __ldg(reinterpret_cast<const float*>(coords) + 4 * n);
  else if (KIND == D3M_COORDS_I64) bf = (float)__ldg(reinterpret_cast<const long long*>(coords) + 4 * n);
  else bf = (float)__ldg(reinterpret_cast<const int*>(coords) + 4 * n);
  if (b < 0) {
    out[n] = make_float4(x, y, z, bf);
    return;
  }
  float gx, gy, gz;
  voxel_world(x, y, z, vs, __ldg(origin + 3 * b), __ldg(origin + 3 * b + 1), __ldg(origin + 3 * b + 2), gx, gy, gz);
  const float4* Wm = reinterpret_cast<const float4*>(w2ac) + 4 * b;
  const float4 r0 = __ldg(Wm), r1 = __ldg(Wm + 1), r2 = __ldg(Wm + 2);
  float4 r;
  r.x = __fmaf_rn(1.0f, r0.w, __fmaf_rn(gz, r0.z, __fmaf_rn(gy, r0.y, __fmul_rn(gx, r0.x))));
  r.y = __fmaf_rn(1.0f, r1.w, __fmaf_rn(gz, r1.z, __fmaf_rn(gy, r1.y, __fmul_rn(gx, r1.x))));
  r.z = __fmaf_rn(1.0f, r2.w, __fmaf_rn(gz, r2.z, __fmaf_rn(gy, r2.y, __fmul_rn(gx, r2.x))));
  r.w = bf;
  out[n] = r;
}

// ------------------------------------------------------------------------------------------------------------------
// get_target (neucon_network.py:52-65): look the GT tsdf / occupancy up at coords // 2^scale in (B,X,Y,Z) volumes.
// Rows outside the volumes (torch would raise IndexError) are reported through `bad` and left untouched.
// ------------------------------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256) gather_targets_kernel(const void* __restrict__ coords, int64_t N, int div,
                                                             const float* __restrict__ tsdf_vol,
                                                             const uint8_t* __restrict__ occ_vol, int B, int X, int Y,
                                                             int Z, float* __restrict__ tsdf_out,
                                                             uint8_t* __restrict__ occ_out, int* __restrict__ bad) {
  pdl_enter();
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  long long b, x, y, z;
  if (KIND == D3M_COORDS_F32) {
    const float4 c = __ldg(reinterpret_cast<const float4*>(coords) + n);
    const float d = (float)div;
    b = (long long)c.x;  // .long() truncates
    x = (long long)floorf(__fdiv_rn(c.y, d));
    y = (long long)floorf(__fdiv_rn(c.z, d));
    z = (long long)floorf(__fdiv_rn(c.w, d));
  } else if (KIND == D3M_COORDS_I64) {
    const longlong2 a = __ldg(reinterpret_cast<const longlong2*>(coords) + 2 * n);
    const longlong2 c = __ldg(reinterpret_cast<const longlong2*>(coords) + 2 * n + 1);
    b = a.x;
    // python floor division
    x = (a.y >= 0) ? a.y / div : -((-a.y + div - 1) / div);
    y = (c.x >= 0) ? c.x / div : -((-c.x + div - 1) / div);
    z = (c.y >= 0) ? c.y / div : -((-c.y + div - 1) / div);
  } else {
    const int4 c = __ldg(reinterpret_cast<const int4*>(coords) + n);
    b = c.x;
    x = (c.y >= 0) ? c.y / div : -((-c.y + div - 1) / div);
    y = (c.z >= 0) ? c.z / div : -((-c.z + div - 1) / div);
    z = (c.w >= 0) ? c.w / div : -((-c.w + div - 1) / div);
  }
  // torch advanced indexing wraps negative indices once
  if (b < 0) b += B;
  if (x < 0) x += X;
  if (y < 0) y += Y;
  if (z < 0) z += Z;
  if (b < 0 || b >= B || x < 0 || x >= X || y < 0 || y >= Y || z < 0 || z >= Z) {
    atomicAdd(bad, 1);
    return;
  }
  const int64_t lin = ((b * X + x) * Y + y) * (int64_t)Z + z;
  if (tsdf_out) tsdf_out[n] = __ldg(tsdf_vol + lin);
  if (occ_out) occ_out[n] = __ldg(occ_vol + lin) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------------------------
// occupancy of the next level (neucon_network.py:181-182):  occ > threshold  and  grid_mask
// grid_mask is either a bool/uint8 array or derived from back_project's count as count > min_count (:132).
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) occupancy_flags_kernel(const float* __restrict__ occ, int64_t occ_stride,
                                                              const float* __restrict__ count, float min_count,
                                                              const uint8_t* __restrict__ mask, float thr, int64_t N,
                                                              uint8_t* __restrict__ flags) {
  pdl_enter();
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  bool f = __ldg(occ + n * occ_stride) > thr;
  if (count) f = f && (__ldg(count + n) > min_count);
  if (mask) f = f && (__ldg(mask + n) != 0);
  flags[n] = f ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------------------------
// Ordered stream compaction = torch.nonzero / boolean-mask indexing (neucon_network.py:192-196, gru_fusion.py).
// Three launches, all integer, output order = input order (deterministic):
//   count  : CTA b counts the set flags of its 4096-item chunk
//   scan   : one CTA turns the chunk counts into exclusive offsets and writes the total
//   write  : CTA b ranks its set flags (warp-shuffle + smem scan) and stores positions (or values[position])
// ------------------------------------------------------------------------------------------------------------------
constexpr int CMP_THREADS = 256;
constexpr int CMP_ITEMS = 16;
constexpr int CMP_CHUNK = CMP_THREADS * CMP_ITEMS;

__device__ __forceinline__ unsigned load_flags16(const uint8_t* __restrict__ flags, int64_t base, int64_t N,
                                                 bool aligned, bool invert) {
  // -> 16-bit mask, bit k = (flags[base+k] != 0) ^ invert, 0 beyond N
  unsigned m = 0;
  if (aligned && base + 16 <= N) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(flags + base));
    const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if ((w[q] >> (8 * k)) & 0xffu) m |= 1u << (4 * q + k);
  } else {
    for (int k = 0; k < 16; ++k)
      if (base + k < N && flags[base + k]) m |= 1u << k;
  }
  if (invert) {
    const int64_t left = N - base;
    m = ~m & (left >= 16 ? 0xffffu : ((1u << (int)left) - 1u));
  }
  return m;
}

__global__ void __launch_bounds__(CMP_THREADS) compact_count_kernel(const uint8_t* __restrict__ flags, int64_t N,
                                                                    bool aligned, bool invert,
                                                                    int64_t* __restrict__ chunk_count) {
  pdl_enter();
  const int64_t base = (int64_t)blockIdx.x * CMP_CHUNK + (int64_t)threadIdx.x * CMP_ITEMS;
  int c = base < N ? __popc(load_flags16(flags, base, N, aligned, invert)) : 0;
  __shared__ int warp_sum[CMP_THREADS / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
#pragma unroll
    for (int w = 0; w < CMP_THREADS / 32; ++w) s += warp_sum[w];
    chunk_count[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(1024) compact_scan_kernel(int64_t* __restrict__ chunk_count, int64_t n_chunks,
                                                            int64_t* __restrict__ total_out) {
  pdl_enter();
  __shared__ int64_t warp_tot[32];
  __shared__ int64_t tile_total;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int64_t carry = 0;  // identical in every thread
  for (int64_t base = 0; base < n_chunks; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const int64_t v = i < n_chunks ? chunk_count[i] : 0;
    int64_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t u = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += u;
    }
    if (lane == 31) warp_tot[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      const int64_t w = warp_tot[lane];
      int64_t winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int64_t u = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += u;
      }
      warp_tot[lane] = winc - w;  // exclusive prefix of each warp
      if (lane == 31) tile_total = winc;
    }
    __syncthreads();
    if (i < n_chunks) chunk_count[i] = carry + warp_tot[wid] + (inc - v);
    carry += tile_total;
    __syncthreads();  // warp_tot / tile_total are rewritten by the next tile
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

template <typename OutT>
__global__ void __launch_bounds__(CMP_THREADS) compact_write_kernel(const uint8_t* __restrict__ flags, int64_t N,
                                                                    bool aligned, bool invert,
                                                                    const int64_t* __restrict__ chunk_offset,
                                                                    const int64_t* __restrict__ values,
                                                                    OutT* __restrict__ out) {
  pdl_enter();
  const int64_t base = (int64_t)blockIdx.x * CMP_CHUNK + (int64_t)threadIdx.x * CMP_ITEMS;
  const unsigned m = base < N ? load_flags16(flags, base, N, aligned, invert) : 0u;
  const int c = __popc(m);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += u;
  }
  __shared__ int warp_tot[CMP_THREADS / 32];
  if (lane == 31) warp_tot[wid] = inc;
  __syncthreads();
  int wbase = 0;
#pragma unroll
  for (int w = 0; w < CMP_THREADS / 32; ++w)
    if (w < wid) wbase += warp_tot[w];
  int64_t pos = __ldg(chunk_offset + blockIdx.x) + wbase + (inc - c);
  unsigned mm = m;
  while (mm) {
    const int k = __ffs(mm) - 1;
    mm &= mm - 1;
    const int64_t i = base + k;
    out[pos++] = (OutT)(values ? __ldg(values + i) : i);
  }
}

// keep[r] = 0 for every r in choice  (neucon_network.py:190-194: occupancy[ind[choice]] = False)
__global__ void __launch_bounds__(256) drop_ranks_kernel(const int64_t* __restrict__ choice, int64_t n_choice,
                                                         int64_t n_keep, uint8_t* __restrict__ keep,
                                                         int* __restrict__ bad) {
  pdl_enter();
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_choice) return;
  const int64_t r = __ldg(choice + j);
  if (r < 0 || r >= n_keep) {
    atomicAdd(bad, 1);
    return;
  }
  keep[r] = 0;
}

// ------------------------------------------------------------------------------------------------------------------
// Row gathers:  dst[m] = src[ind[m]]  (boolean-mask indexing after compaction), rows of `words` 32-bit words, and the
// fused  cat([a[ind], b[ind], c[ind], d[ind]], dim=1)  of neucon_network.py:203-207.
// ------------------------------------------------------------------------------------------------------------------
template <typename VecT>
__global__ void __launch_bounds__(256) gather_rows_kernel(const VecT* __restrict__ src, int vec_per_row,
                                                          const int64_t* __restrict__ ind, int64_t total,
                                                          VecT* __restrict__ dst) {
  pdl_enter();
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int64_t m = t / vec_per_row;
  const int j = (int)(t - m * vec_per_row);
  dst[t] = __ldg(src + __ldg(ind + m) * vec_per_row + j);
}

struct ConcatSrc {
  const float* p[4];
  int w[4];       // row width of each source in floats
  int begin[5];   // column where each source starts in the output row
  int n;
};

// Generic path (any widths / alignment): one thread per output float.
__global__ void __launch_bounds__(256) gather_concat_kernel(ConcatSrc s, const int64_t* __restrict__ ind,
                                                            int64_t total, float* __restrict__ dst) {
  pdl_enter();
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= total) return;
  const int wtot = s.begin[s.n];
  const int64_t m = t / wtot;
  const int col = (int)(t - m * wtot);
  const int64_t r = ind ? __ldg(ind + m) : m;
  int k = 0;
#pragma unroll
  for (int q = 1; q < 4; ++q)
    if (q < s.n && col >= s.begin[q]) k = q;
  dst[t] = __ldg(s.p[k] + r * s.w[k] + (col - s.begin[k]));
}

// Pair path (even row width, 8-byte aligned destination -- the NeuralRecon case: C + 2 floats with C = 96/48/24):
// one thread = GC_UNROLL 64-bit units of the flat output; a unit never straddles a row because the width is even.
// Units that lie inside one even-aligned source are fetched with one 64-bit load, the (tsdf, occ) tail pair with two
// 32-bit loads.  No shared memory, no barrier: all of a thread's loads are in flight before its first store, and a
// warp's stores are one contiguous 256-byte span.
constexpr int GC_UNROLL = 4;
__global__ void __launch_bounds__(256) gather_concat_pair_kernel(ConcatSrc s, const int64_t* __restrict__ ind,
                                                                 uint32_t total_units, uint32_t units_per_row,
                                                                 float2* __restrict__ dst) {
  pdl_enter();
  const uint32_t t0 = (blockIdx.x * (uint32_t)blockDim.x * GC_UNROLL) + threadIdx.x;
  float2 v[GC_UNROLL];
#pragma unroll
  for (int q = 0; q < GC_UNROLL; ++q) {
    const uint32_t t = t0 + q * blockDim.x;
    v[q] = make_float2(0.f, 0.f);
    if (t < total_units) {
      const uint32_t m = t / units_per_row;
      const int col = (int)(t - m * units_per_row) * 2;
      const int64_t r = ind ? __ldg(ind + m) : (int64_t)m;
      int k = 0;
#pragma unroll
      for (int j = 1; j < 4; ++j)
        if (j < s.n && col >= s.begin[j]) k = j;
      const int c0 = col - s.begin[k];
      const float* p0 = s.p[k] + r * s.w[k] + c0;
      if (c0 + 1 < s.w[k]) {
        if ((reinterpret_cast<uintptr_t>(p0) & 7u) == 0) {
          v[q] = __ldg(reinterpret_cast<const float2*>(p0));
        } else {
          v[q] = make_float2(__ldg(p0), __ldg(p0 + 1));
        }
      } else {
        int k1 = k + 1;
        while (k1 < s.n - 1 && s.w[k1] == 0) ++k1;   // col+1 < row width, so a non-empty source follows
        v[q] = make_float2(__ldg(p0), __ldg(s.p[k1] + r * s.w[k1]));
      }
    }
  }
#pragma unroll
  for (int q = 0; q < GC_UNROLL; ++q) {
    const uint32_t t = t0 + q * blockDim.x;
    if (t < total_units) __stcs(dst + t, v[q]);
  }
}

// Tiled path: a CTA assembles GC_ROWS consecutive OUTPUT rows in shared memory -- every source row is fetched with the
// widest vector its width and base alignment allow (128-bit for the 24/48/96-float feature rows), the index is read
// once per row -- and then streams the tile, which is one contiguous span of GC_ROWS*wtot floats starting on a
// 128-byte boundary, to global memory with 128-bit stores.  Output rows are C+2 floats wide (not a multiple of 16
// bytes), which is why a direct row-wise vector store is impossible and the one-float-per-thread version ran at 13 %
// of the HBM peak.
constexpr int GC_ROWS = 32;
__global__ void __launch_bounds__(256) gather_concat_tile_kernel(ConcatSrc s, const int64_t* __restrict__ ind, int64_t M,
                                                                 float* __restrict__ dst) {
  pdl_enter();
  extern __shared__ __align__(16) float gc_tile[];
  __shared__ int64_t rows[GC_ROWS];
  const int wtot = s.begin[s.n];
  const int64_t m0 = (int64_t)blockIdx.x * GC_ROWS;
  const int nrow = (int)min((int64_t)GC_ROWS, M - m0);
  if (threadIdx.x < nrow) rows[threadIdx.x] = ind ? __ldg(ind + m0 + threadIdx.x) : m0 + threadIdx.x;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (k >= s.n || s.w[k] == 0) continue;
    const int w = s.w[k];
    const float* __restrict__ src = s.p[k];
    if ((w & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
      const int vpr = w >> 2;
      for (int t = threadIdx.x; t < nrow * vpr; t += blockDim.x) {
        const int r = t / vpr, j = t - r * vpr;
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + rows[r] * w) + j);
        float* o = gc_tile + r * wtot + s.begin[k] + 4 * j;
        o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
      }
    } else {
      for (int t = threadIdx.x; t < nrow * w; t += blockDim.x) {
        const int r = t / w, j = t - r * w;
        gc_tile[r * wtot + s.begin[k] + j] = __ldg(src + rows[r] * w + j);
      }
    }
  }
  __syncthreads();
  float* out = dst + m0 * wtot;
  const int nflt = nrow * wtot;
  if ((reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
    const int nvec = nflt >> 2;
    for (int t = threadIdx.x; t < nvec; t += blockDim.x)
      __stcs(reinterpret_cast<float4*>(out) + t, reinterpret_cast<const float4*>(gc_tile)[t]);
    for (int t = (nvec << 2) + threadIdx.x; t < nflt; t += blockDim.x) out[t] = gc_tile[t];
  } else {
    for (int t = threadIdx.x; t < nflt; t += blockDim.x) out[t] = gc_tile[t];
  }
}

// number of rows per fragment (neucon_network.py:197-201 "no valid points: scale, batch")
template <int KIND>
__global__ void __launch_bounds__(256) batch_counts_kernel(const void* __restrict__ coords, int64_t N, int B,
                                                           unsigned long long* __restrict__ counts) {
  pdl_enter();
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int b = -1;
  if (n < N) {
    float x, y, z;
    b = load_coord<KIND>(coords, n, B, x, y, z);
  }
  // warp-aggregate when the whole warp agrees (the common case: rows are fragment-sorted)
  const unsigned act = __activemask();
  const int b0 = __shfl_sync(act, b, __ffs(act) - 1);
  const unsigned same = __ballot_sync(act, b == b0);
  if (same == act) {
    if (b >= 0 && (threadIdx.x & 31) == __ffs(act) - 1) atomicAdd(counts + b, (unsigned long long)__popc(act));
  } else if (b >= 0) {
    atomicAdd(counts + b, 1ull);
  }
}

#define D3M_DISPATCH_KIND(kind, CALL)                                      \
  switch (kind) {                                                          \
    case D3M_COORDS_F32: { constexpr int K = D3M_COORDS_F32; CALL; } break; \
    case D3M_COORDS_I64: { constexpr int K = D3M_COORDS_I64; CALL; } break; \
    default: { constexpr int K = D3M_COORDS_I32; CALL; } break;            \
  }

static int check_kind(int kind, const void* p, const char* what) {
  D3M_REQUIRE(kind == D3M_COORDS_F32 || kind == D3M_COORDS_I64 || kind == D3M_COORDS_I32, D3M_ERR_ARG,
              "%s: unknown coords_kind %d", what, kind);
  D3M_REQUIRE(aligned16(p), D3M_ERR_ALIGN, "%s: coords must be 16-byte aligned", what);
  return D3M_OK;
}

#define D3M_NEED_DEVICE(what) \
  D3M_REQUIRE(d3m_device_count() > 0, D3M_ERR_NO_DEVICE, what ": no CUDA device (there is no CPU fallback)")

}  // namespace d3m

using namespace d3m;

extern "C" int d3m_grid_coords(int nx, int ny, int nz, int interval, int B, float* coords, float* grid3,
                               void* stream_) {
  D3M_NEED_DEVICE("d3m_grid_coords");
  D3M_REQUIRE(nx >= 0 && ny >= 0 && nz >= 0 && interval >= 1 && B >= 0 && (coords || grid3), D3M_ERR_ARG,
              "d3m_grid_coords: bad arguments");
  D3M_REQUIRE(!coords || aligned16(coords), D3M_ERR_ALIGN, "d3m_grid_coords: coords must be 16-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int gx = (nx + interval - 1) / interval, gy = (ny + interval - 1) / interval, gz = (nz + interval - 1) / interval;
  const int64_t n_per = (int64_t)gx * gy * gz;
  const int64_t total = n_per * (coords ? B : 1);
  if (total == 0) return D3M_OK;
  LaunchScope ls("grid_coords", stream);
  launch_k(grid_coords_kernel, dim3(blocks_for(total, 256)), dim3(256), 0, stream, gy, gz, n_per, interval, B,
                                                                reinterpret_cast<float4*>(coords), grid3);
  D3M_CUDA_CHECK(cudaGetLastError());
  return D3M_OK;
}

extern "C" int d3m_upsample(const void* pre_coords, int coords_kind, const float* pre_feat, int64_t N, int C,
                            int interval, int num, void* up_coords, float* up_feat, void* stream_) {
  D3M_NEED_DEVICE("d3m_upsample");
  D3M_REQUIRE(N >= 0 && num >= 1 && num <= 8 && C >= 0, D3M_ERR_ARG, "d3m_upsample: bad arguments (num must be 1..8)");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (N == 0) return D3M_OK;
  if (pre_coords) {
    D3M_REQUIRE(up_coords != nullptr, D3M_ERR_ARG, "d3m_upsample: up_coords is NULL");
    int rc = check_kind(coords_kind, pre_coords, "d3m_upsample");
    if (rc) return rc;
    D3M_REQUIRE(aligned16(up_coords), D3M_ERR_ALIGN, "d3m_upsample: up_coords must be 16-byte aligned");
    const int64_t total = N * num;
    LaunchScope ls("upsample_coords", stream);
    D3M_DISPATCH_KIND(coords_kind, (launch_k(upsample_coords_kernel<K>, dim3(blocks_for(total, 256)), dim3(256), 0, stream, 
                                       pre_coords, total, num, interval, up_coords)));
    D3M_CUDA_CHECK(cudaGetLastError());
  }
  if (pre_feat && C > 0) {
    D3M_REQUIRE(up_feat != nullptr, D3M_ERR_ARG, "d3m_upsample: up_feat is NULL");
    LaunchScope ls("upsample_feat", stream);
    const uintptr_t al = reinterpret_cast<uintptr_t>(pre_feat) | reinterpret_cast<uintptr_t>(up_feat);
    if (C % 4 == 0 && (al & 15u) == 0) {
      launch_upsample_feat<float4>(pre_feat, N, C, num, up_feat, stream);
    } else if (C % 2 == 0 && (al & 7u) == 0) {
      launch_upsample_feat<float2>(pre_feat, N, C, num, up_feat, stream);
    } else {
      launch_upsample_feat<float>(pre_feat, N, C, num, up_feat, stream);
    }
    D3M_CUDA_CHECK(cudaGetLastError());
  }
  return D3M_OK;
}

extern "C" int d3m_aligned_camera_coords(const void* coords, int coords_kind, int64_t N, const float* origin, int B,
                                         float voxel_size, const float* world_to_aligned_camera, float* r_coords,
                                         void* stream_) {
  D3M_NEED_DEVICE("d3m_aligned_camera_coords");
  D3M_REQUIRE(N >= 0 && B >= 0, D3M_ERR_ARG, "d3m_aligned_camera_coords: bad arguments");
  if (N == 0) return D3M_OK;
  D3M_REQUIRE(coords && r_coords && (B == 0 || (origin && world_to_aligned_camera)), D3M_ERR_ARG,
              "d3m_aligned_camera_coords: NULL pointer");
  int rc = check_kind(coords_kind, coords, "d3m_aligned_camera_coords");
  if (rc) return rc;
  D3M_REQUIRE(aligned16(r_coords) && aligned16(world_to_aligned_camera), D3M_ERR_ALIGN,
              "d3m_aligned_camera_coords: r_coords / matrices must be 16-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LaunchScope ls("aligned_camera_coords", stream);
  D3M_DISPATCH_KIND(coords_kind, (launch_k(aligned_camera_kernel<K>, dim3(blocks_for(N, 256)), dim3(256), 0, stream, 
                                     coords, N, origin, B, voxel_size, world_to_aligned_camera,
                                     reinterpret_cast<float4*>(r_coords))));
  D3M_CUDA_CHECK(cudaGetLastError());
  return D3M_OK;
}

extern "C" int d3m_gather_targets(const void* coords, int coords_kind, int64_t N, int divisor, const float* tsdf_vol,
                                  const uint8_t* occ_vol, int B, int X, int Y, int Z, float* tsdf_out,
                                  uint8_t* occ_out, int* bad_rows, void* stream_) {
  D3M_NEED_DEVICE("d3m_gather_targets");
  D3M_REQUIRE(N >= 0 && divisor >= 1 && B >= 0 && X >= 0 && Y >= 0 && Z >= 0, D3M_ERR_ARG,
              "d3m_gather_targets: bad arguments");
  D3M_REQUIRE(bad_rows != nullptr, D3M_ERR_ARG, "d3m_gather_targets: bad_rows is NULL");
  D3M_REQUIRE((!tsdf_out || tsdf_vol) && (!occ_out || occ_vol), D3M_ERR_ARG, "d3m_gather_targets: volume is NULL");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  D3M_CUDA_CHECK(cudaMemsetAsync(bad_rows, 0, sizeof(int), stream));
  if (N == 0) return D3M_OK;
  int rc = check_kind(coords_kind, coords, "d3m_gather_targets");
  if (rc) return rc;
  LaunchScope ls("gather_targets", stream);
  D3M_DISPATCH_KIND(coords_kind, (launch_k(gather_targets_kernel<K>, dim3(blocks_for(N, 256)), dim3(256), 0, stream, 
                                     coords, N, divisor, tsdf_vol, occ_vol, B, X, Y, Z, tsdf_out, occ_out, bad_rows)));
  D3M_CUDA_CHECK(cudaGetLastError());
  return D3M_OK;
}

extern "C" int d3m_occupancy_flags(const float* occ, int64_t occ_stride, const float* count, float min_count,
                                   const uint8_t* grid_mask, float threshold, int64_t N, uint8_t* flags,
                                   void* stream_) {
  D3M_NEED_DEVICE("d3m_occupancy_flags");
  D3M_REQUIRE(N >= 0 && occ_stride >= 1, D3M_ERR_ARG, "d3m_occupancy_flags: bad arguments");
  if (N == 0) return D3M_OK;
  D3M_REQUIRE(occ && flags, D3M_ERR_ARG, "d3m_occupancy_flags: NULL pointer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LaunchScope ls("occupancy_flags", stream);
  launch_k(occupancy_flags_kernel, dim3(blocks_for(N, 256)), dim3(256), 0, stream, occ, occ_stride, count, min_count, grid_mask,
                                                                threshold, N, flags);
  D3M_CUDA_CHECK(cudaGetLastError());
  return D3M_OK;
}

extern "C" size_t d3m_compact_workspace(int64_t N) {
  const int64_t chunks = N > 0 ? (N + CMP_CHUNK - 1) / CMP_CHUNK : 0;
  return align_up((size_t)(chunks + 1) * sizeof(int64_t), 256);
}

extern "C" int d3m_compact(const uint8_t* flags, int64_t N, int invert, const int64_t* values, int64_t* out,
                           int64_t* total_dev, void* workspace, size_t workspace_bytes, void* stream_) {
  D3M_NEED_DEVICE("d3m_compact");
  D3M_REQUIRE(N >= 0 && total_dev, D3M_ERR_ARG, "d3m_compact: bad arguments");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (N == 0) {
    D3M_CUDA_CHECK(cudaMemsetAsync(total_dev, 0, sizeof(int64_t), stream));
    return D3M_OK;
  }
  D3M_REQUIRE(flags && out && workspace, D3M_ERR_ARG, "d3m_compact: NULL pointer");
  D3M_REQUIRE(workspace_bytes >= d3m_compact_workspace(N), D3M_ERR_WORKSPACE, "d3m_compact: workspace too small");
  D3M_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 7u) == 0, D3M_ERR_ALIGN, "d3m_compact: workspace alignment");
  const int64_t chunks = (N + CMP_CHUNK - 1) / CMP_CHUNK;
  D3M_REQUIRE(chunks < (1ll << 31), D3M_ERR_ARG, "d3m_compact: N too large");
  int64_t* chunk = static_cast<int64_t*>(workspace);
  const bool al = aligned16(flags);
  {
    LaunchScope ls("compact_count", stream);
    launch_k(compact_count_kernel, dim3((unsigned)chunks), dim3(CMP_THREADS), 0, stream, flags, N, al, invert != 0, chunk);
    D3M_CUDA_CHECK(cudaGetLastError());
  }
  {
    LaunchScope ls("compact_scan", stream);
    launch_k(compact_scan_kernel, dim3(1), dim3(1024), 0, stream, chunk, chunks, total_dev);
    D3M_CUDA_CHECK(cudaGetLastError());
  }
  {
    LaunchScope ls("compact_write", stream);
    launch_k(compact_write_kernel<int64_t>, dim3((unsigned)chunks), dim3(CMP_THREADS), 0, stream, flags, N, al, invert != 0, chunk, values,
                                                                                 out);
    D3M_CUDA_CHECK(cudaGetLastError());
  }
  return D3M_OK;
}

extern "C" int d3m_drop_ranks(const int64_t* choice, int64_t n_choice, int64_t n_keep, uint8_t* keep, int* bad,
                              void* stream_) {
  D3M_NEED_DEVICE("d3m_drop_ranks");
  D3M_REQUIRE(n_choice >= 0 && n_keep >= 0 && bad, D3M_ERR_ARG, "d3m_drop_ranks: bad arguments");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  D3M_CUDA_CHECK(cudaMemsetAsync(bad, 0, sizeof(int), stream));
  if (n_keep == 0) return D3M_OK;
  D3M_REQUIRE(keep != nullptr, D3M_ERR_ARG, "d3m_drop_ranks: keep is NULL");
  D3M_CUDA_CHECK(cudaMemsetAsync(keep, 1, (size_t)n_keep, stream));
  if (n_choice == 0) return D3M_OK;
  D3M_REQUIRE(choice != nullptr, D3M_ERR_ARG, "d3m_drop_ranks: choice is NULL");
  LaunchScope ls("drop_ranks", stream);
  launch_k(drop_ranks_kernel, dim3(blocks_for(n_choice, 256)), dim3(256), 0, stream, choice, n_choice, n_keep, keep, bad);
  D3M_CUDA_CHECK(cudaGetLastError());
  return D3M_OK;
}

extern "C" int d3m_gather_rows(const void* src, int64_t row_bytes, const int64_t* ind, int64_t M, void* dst,
                               void* stream_) {
  D3M_NEED_DEVICE("d3m_gather_rows");
  D3M_REQUIRE(M >= 0 && row_bytes >= 0 && row_bytes % 4 == 0 && row_bytes < (1ll << 31), D3M_ERR_ARG,
              "d3m_gather_rows: row_bytes must be a multiple of 4");
  if (M == 0 || row_bytes == 0) return D3M_OK;
  D3M_REQUIRE(src && ind && dst, D3M_ERR_ARG, "d3m_gather_rows: NULL pointer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LaunchScope ls("gather_rows", stream);
  if (row_bytes % 16 == 0 && aligned16(src) && aligned16(dst)) {
    const int v = (int)(row_bytes / 16);
    launch_k(gather_rows_kernel<uint4>, dim3(blocks_for(M * v, 256)), dim3(256), 0, stream, static_cast<const uint4*>(src), v, ind,
                                                                        M * v, static_cast<uint4*>(dst));
  } else {
    const int v = (int)(row_bytes / 4);
    launch_k(gather_rows_kernel<uint32_t>, dim3(blocks_for(M * v, 256)), dim3(256), 0, stream, static_cast<const uint32_t*>(src), v, ind,
                                                                           M * v, static_cast<uint32_t*>(dst));
  }
  D3M_CUDA_CHECK(cudaGetLastError());
  return D3M_OK;
}

extern "C" int d3m_gather_concat(const float* const* srcs_host, const int* widths_host, int n_src, const int64_t* ind,
                                 int64_t M, float* dst, void* stream_) {
  D3M_NEED_DEVICE("d3m_gather_concat");
  D3M_REQUIRE(n_src >= 1 && n_src <= 4 && M >= 0 && srcs_host && widths_host, D3M_ERR_ARG,
              "d3m_gather_concat: 1..4 sources");
  ConcatSrc s;
  s.n = n_src;
  s.begin[0] = 0;
  for (int k = 0; k < 4; ++k) {
    s.p[k] = k < n_src ? srcs_host[k] : nullptr;
    s.w[k] = k < n_src ? widths_host[k] : 0;
    D3M_REQUIRE(s.w[k] >= 0 && (k >= n_src || s.w[k] == 0 || s.p[k]), D3M_ERR_ARG, "d3m_gather_concat: bad source %d", k);
    s.begin[k + 1] = s.begin[k] + s.w[k];
  }
  const int64_t total = M * s.begin[n_src];
  if (total == 0) return D3M_OK;
  D3M_REQUIRE(dst != nullptr, D3M_ERR_ARG, "d3m_gather_concat: dst is NULL");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LaunchScope ls("gather_concat", stream);
  const size_t tile_bytes = (size_t)GC_ROWS * s.begin[n_src] * sizeof(float);
  const int wtot = s.begin[n_src];
  if (wtot % 2 == 0 && (reinterpret_cast<uintptr_t>(dst) & 7u) == 0 && total / 2 < (1ll << 32) - (1 << 20)) {
    const int64_t units = total / 2;
    launch_k(gather_concat_pair_kernel, dim3(blocks_for(units, 256 * GC_UNROLL)), dim3(256), 0, stream, 
        s, ind, (uint32_t)units, (uint32_t)(wtot / 2), reinterpret_cast<float2*>(dst));
  } else if (tile_bytes <= 40 * 1024) {
    launch_k(gather_concat_tile_kernel, dim3(blocks_for(M, GC_ROWS)), dim3(256), tile_bytes, stream, s, ind, M, dst);
  } else {
    launch_k(gather_concat_kernel, dim3(blocks_for(total, 256)), dim3(256), 0, stream, s, ind, total, dst);
  }
  D3M_CUDA_CHECK(cudaGetLastError());
  return D3M_OK;
}

extern "C" int d3m_batch_counts(const void* coords, int coords_kind, int64_t N, int B, int64_t* counts,
                                void* stream_) {
  D3M_NEED_DEVICE("d3m_batch_counts");
  D3M_REQUIRE(N >= 0 && B >= 0 && (B == 0 || counts), D3M_ERR_ARG, "d3m_batch_counts: bad arguments");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (B == 0) return D3M_OK;
  D3M_CUDA_CHECK(cudaMemsetAsync(counts, 0, sizeof(int64_t) * B, stream));
  if (N == 0) return D3M_OK;
  int rc = check_kind(coords_kind, coords, "d3m_batch_counts");
  if (rc) return rc;
  LaunchScope ls("batch_counts", stream);
  D3M_DISPATCH_KIND(coords_kind, (launch_k(batch_counts_kernel<K>, dim3(blocks_for(N, 256)), dim3(256), 0, stream, 
                                     coords, N, B, reinterpret_cast<unsigned long long*>(counts))));
  D3M_CUDA_CHECK(cudaGetLastError());
  return D3M_OK;
}
