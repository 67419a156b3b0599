// TSDF fusion for sm_100a.  Replaces the PyCUDA `integrate` kernel + host loop of the reference
// (deep3dmap/core/tsdf/tsdf_volume.py:68-126, 210-256) and the torch-CPU `integrate()` used by the
// dataloader (tsdf_volume.py:437-482).
//
// Design: the reference launches one thread per voxel of the WHOLE volume for every frame (134 M threads at
// 512^3, >99 % of which exit at the frustum test) and copies seven arrays both ways per call.  Here
//   * a tiny prep kernel reduces each frame's max depth (16 partial maxima per frame),
//   * ONE persistent kernel handles any number of frames: every CTA derives the union of the frames' frustum
//     boxes, walks the 8x8x16-voxel tiles inside it, culls frames per tile against the frustum planes,
//     keeps the tile's tsdf/weight values in REGISTERS while it applies the surviving frames in order (the
//     running average is order dependent), and writes back only what changed.  Volume bytes move once per
//     launch instead of once per frame; z-fastest rows give coalesced 64-byte segments.
// Per-voxel arithmetic is the reference's, rounding step for rounding step (explicit _rn intrinsics
// reproducing the FMA contraction nvcc applies to the reference source; checked bit-for-bit against that
// source compiled verbatim, see oracle/build_ref.py and tests/test_gpu_tsdf.py).
// Deliberate deviation: integer index decomposition (the reference's float one, :89-91, breaks for > 2^24 voxels).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <condition_variable>
#include <map>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "d3m_common.cuh"

namespace d3m {

constexpr int kTileX = 8, kTileY = 8, kTileZ = 16;
constexpr int kTsdfThreads = 256;         // 8 warps = 2x2x2 sub-boxes of 4x4x8 voxels; a lane owns 4 x-planes of one (y,z)
constexpr int kVoxPerThread = 4;
constexpr int kMaxFramesPerLaunch = 1024;
constexpr int kPrepParts = 16;
constexpr int kRing = 4;
constexpr float kSlack = 0.02f;           // metres; conservative margin of the cull tests

struct Frame {
  float fx, fy, cx, cy;
  float T[12];       // kernel semantics: rows 0..2 of cam->world pose; torch semantics: rows 0..2 of world->cam
  float obs;
  float centre[3];   // camera centre, world
  float dirs[4][3];  // world-space rays through the 4 image corners (per unit camera depth)
  float planes[5][4];  // near, left, right, top, bottom: unit normal (world) and offset; inside if n.p + d >= 0
  float far_n[3];
  float far_d0;        // inside if far_n.p + far_d0 + zmax >= 0
};

struct TsdfParams {
  float* tsdf;
  float* weight;
  float* color;
  int dx, dy, dz;
  int xoff;             // global x index of local plane 0 (slab sharding), 0 otherwise
  int variant;          // tuning switches (D3M_TSDF_VARIANT, default 0): 1 = warp-level frame cull, 2 = division-free
                        // quick reject.  Both were measured SLOWER on B200 (profiles/r01_tsdf_variants.txt): the extra
                        // tests cost more issue slots than the divisions they save
  float ox, oy, oz, vs, trunc;
  const Frame* frames;
  int F;
  const float* depth;   // (F,H,W)
  const float* cimg;    // (F,H,W) folded colour or NULL
  int H, W;
  float* partial_max;   // (F, kPrepParts)
};

__global__ void __launch_bounds__(256) tsdf_prep_kernel(const float* __restrict__ depth, int HW,
                                                        float* __restrict__ partial_max) {
  __shared__ float red[8];
  const int f = blockIdx.y, part = blockIdx.x;
  const float* d = depth + (int64_t)f * HW;
  const int per = (HW + kPrepParts - 1) / kPrepParts;
  const int i0 = part * per, i1 = min(HW, i0 + per);
  float m = 0.0f;
  for (int i = i0 + threadIdx.x; i < i1; i += 256) {
    // a NaN depth passes the reference's `depth == 0` / `diff < -trunc` tests (tsdf_volume.py:114,119), so it can
    // update voxels at any distance: disable the far cull for such a frame instead of ignoring the pixel
    const float v = __ldg(d + i);
    m = (v != v) ? INFINITY : fmaxf(m, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
    partial_max[f * kPrepParts + part] = m;
  }
}

__device__ __forceinline__ float frame_zmax(const TsdfParams& p, int f) {
  float m = 0.0f;
#pragma unroll
  for (int i = 0; i < kPrepParts; ++i) m = fmaxf(m, p.partial_max[f * kPrepParts + i]);
  return m;  // 0 -> frame has no valid depth
}

// frustum box of frame f in tile coordinates (inclusive), false when empty
__device__ __forceinline__ bool frame_tile_box(const TsdfParams& p, const Frame& fr, float zmaxd, int lo[3], int hi[3]) {
  if (!(zmaxd > 0.0f)) return false;
  const float zm = zmaxd + p.trunc + kSlack;
  float mn[3], mx[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) { mn[a] = fr.centre[a]; mx[a] = fr.centre[a]; }
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float q = fr.centre[a] + zm * fr.dirs[k][a];
      mn[a] = fminf(mn[a], q);
      mx[a] = fmaxf(mx[a], q);
    }
  const float org[3] = {p.ox + (float)p.xoff * p.vs, p.oy, p.oz};  // cull geometry only (conservative)
  const int dims[3] = {p.dx, p.dy, p.dz};
  const int tdim[3] = {kTileX, kTileY, kTileZ};
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float vlo = floorf((mn[a] - kSlack - org[a]) / p.vs) - 1.0f;
    const float vhi = ceilf((mx[a] + kSlack - org[a]) / p.vs) + 1.0f;
    if (vhi < 0.0f || vlo > (float)(dims[a] - 1)) return false;
    const int ilo = (int)fmaxf(vlo, 0.0f), ihi = (int)fminf(vhi, (float)(dims[a] - 1));
    lo[a] = ilo / tdim[a];
    hi[a] = ihi / tdim[a];
  }
  return true;
}

__device__ __forceinline__ bool tile_hits_frame(const Frame& fr, float zmaxd, float trunc, const float c[3],
                                                const float h[3]) {
  if (!(zmaxd > 0.0f)) return false;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const float* pl = fr.planes[k];
    const float dist = pl[0] * c[0] + pl[1] * c[1] + pl[2] * c[2] + pl[3];
    const float reach = fabsf(pl[0]) * h[0] + fabsf(pl[1]) * h[1] + fabsf(pl[2]) * h[2];
    if (dist + reach < -kSlack) return false;
  }
  const float dist = fr.far_n[0] * c[0] + fr.far_n[1] * c[1] + fr.far_n[2] * c[2] + fr.far_d0 + zmaxd + trunc;
  const float reach = fabsf(fr.far_n[0]) * h[0] + fabsf(fr.far_n[1]) * h[1] + fabsf(fr.far_n[2]) * h[2];
  return dist + reach >= -kSlack;
}

// PTX cvt.rzi.s32.f32 is what `(int)` compiles to: saturating, NaN -> 0 (same as in the reference kernel)
__device__ __forceinline__ int f2i_rz(float x) { return __float2int_rz(x); }

// Division-free conservative rejection: true only for voxels the exact arithmetic below certainly skips -- behind the
// camera, a full pixel outside the image (the exact test rounds to the nearest pixel, i.e. has a half-pixel band), or
// so far behind the deepest measurement of the frame that depth - cam_z < -trunc for every pixel.  Voxels near any
// boundary fall through to the exact reference arithmetic, so results are unchanged bit for bit.
__device__ __forceinline__ bool quick_reject(const Frame& fr, float camx, float camy, float camz, float zfar, float W,
                                             float H) {
  if (camz < 0.0f || camz > zfar) return true;
  const float ax = fr.fx * camx, ay = fr.fy * camy;
  if (ax + (fr.cx + 1.0f) * camz < 0.0f) return true;   // u < -1
  if (ax + (fr.cx - W) * camz > 0.0f) return true;      // u > W
  if (ay + (fr.cy + 1.0f) * camz < 0.0f) return true;   // v < -1
  if (ay + (fr.cy - H) * camz > 0.0f) return true;      // v > H
  return false;
}

template <int SEM, bool COLOR>
__device__ __forceinline__ void integrate_voxel(const TsdfParams& p, const Frame& fr, const float* __restrict__ depth,
                                                const float* __restrict__ cimg, float zfar, float vx, float vy,
                                                float vz, float& tsdf, float& w, float& col, bool& dirty) {
  float camx, camy, camz;
  int px, py;
  bool ok;
  if (SEM == D3M_TSDF_KERNEL_SEMANTICS) {
    // tsdf_volume.py:94-106 with the contraction of the reference build
    const float ptx = __fmaf_rn(vx, p.vs, p.ox), pty = __fmaf_rn(vy, p.vs, p.oy), ptz = __fmaf_rn(p.vs, vz, p.oz);
    const float tx = __fsub_rn(ptx, fr.T[3]), ty = __fsub_rn(pty, fr.T[7]), tz = __fsub_rn(ptz, fr.T[11]);
    camx = __fmaf_rn(tz, fr.T[8], __fmaf_rn(tx, fr.T[0], __fmul_rn(ty, fr.T[4])));
    camy = __fmaf_rn(tz, fr.T[9], __fmaf_rn(tx, fr.T[1], __fmul_rn(ty, fr.T[5])));
    camz = __fmaf_rn(tz, fr.T[10], __fmaf_rn(tx, fr.T[2], __fmul_rn(ty, fr.T[6])));
    if ((p.variant & 2) && quick_reject(fr, camx, camy, camz, zfar, (float)p.W, (float)p.H)) return;
    px = f2i_rz(roundf(__fmaf_rn(fr.fx, __fdiv_rn(camx, camz), fr.cx)));
    py = f2i_rz(roundf(__fmaf_rn(fr.fy, __fdiv_rn(camy, camz), fr.cy)));
    ok = !(px < 0 || px >= p.W || py < 0 || py >= p.H || camz < 0.0f);  // :110
  } else {
    // tsdf_volume.py:523 (world_c), :451-459
    const float wx = __fadd_rn(p.ox, __fmul_rn(p.vs, vx)), wy = __fadd_rn(p.oy, __fmul_rn(p.vs, vy)),
                wz = __fadd_rn(p.oz, __fmul_rn(p.vs, vz));
    camx = __fadd_rn(__fmaf_rn(fr.T[2], wz, __fmaf_rn(fr.T[1], wy, __fmul_rn(fr.T[0], wx))), fr.T[3]);
    camy = __fadd_rn(__fmaf_rn(fr.T[6], wz, __fmaf_rn(fr.T[5], wy, __fmul_rn(fr.T[4], wx))), fr.T[7]);
    camz = __fadd_rn(__fmaf_rn(fr.T[10], wz, __fmaf_rn(fr.T[9], wy, __fmul_rn(fr.T[8], wx))), fr.T[11]);
    if ((p.variant & 2) && quick_reject(fr, camx, camy, camz, zfar, (float)p.W, (float)p.H)) return;
    const float rx = rintf(__fadd_rn(__fdiv_rn(__fmul_rn(camx, fr.fx), camz), fr.cx));
    const float ry = rintf(__fadd_rn(__fdiv_rn(__fmul_rn(camy, fr.fy), camz), fr.cy));
    ok = (rx >= 0.0f) && (rx < (float)p.W) && (ry >= 0.0f) && (ry < (float)p.H) && (camz > 0.0f);  // :462
    px = ok ? (int)rx : 0;
    py = ok ? (int)ry : 0;
  }
  if (!ok) return;
  const float d = __ldg(depth + py * p.W + px);
  const float diff = __fsub_rn(d, camz);
  if (SEM == D3M_TSDF_KERNEL_SEMANTICS) {
    if (d == 0.0f) return;            // :114
    if (diff < -p.trunc) return;      // :119
    const float dist = fminf(__fdiv_rn(diff, p.trunc), 1.0f);
    const float w_old = w;
    const float w_new = __fadd_rn(w_old, fr.obs);
    w = w_new;
    tsdf = __fdiv_rn(__fmaf_rn(dist, fr.obs, __fmul_rn(w_old, tsdf)), w_new);  // :121-126
    dirty = true;
    if (COLOR) {
      // :130-141 (unreachable in the reference because of the `return` at :129); all values are integers < 2^24
      const float oc = col;
      const float ob = floorf(__fmul_rn(oc, 1.0f / 65536.0f));
      const float t0 = __fsub_rn(oc, __fmul_rn(ob, 65536.0f));
      const float og = floorf(__fmul_rn(t0, 1.0f / 256.0f));
      const float orr = __fsub_rn(t0, __fmul_rn(og, 256.0f));
      const float nc = __ldg(cimg + py * p.W + px);
      float nb = floorf(__fmul_rn(nc, 1.0f / 65536.0f));
      const float t1 = __fsub_rn(nc, __fmul_rn(nb, 65536.0f));
      float ng = floorf(__fmul_rn(t1, 1.0f / 256.0f));
      float nr = __fsub_rn(t1, __fmul_rn(ng, 256.0f));
      nb = fminf(roundf(__fdiv_rn(__fmaf_rn(w_old, ob, __fmul_rn(fr.obs, nb)), w_new)), 255.0f);
      ng = fminf(roundf(__fdiv_rn(__fmaf_rn(w_old, og, __fmul_rn(fr.obs, ng)), w_new)), 255.0f);
      nr = fminf(roundf(__fdiv_rn(__fmaf_rn(w_old, orr, __fmul_rn(fr.obs, nr)), w_new)), 255.0f);
      col = __fadd_rn(__fadd_rn(__fmul_rn(nb, 65536.0f), __fmul_rn(ng, 256.0f)), nr);
    }
  } else {
    if (!(d > 0.0f && diff >= -p.trunc)) return;  // :471
    float dist = __fdiv_rn(diff, p.trunc);
    if (dist > 1.0f) dist = 1.0f;                 // clamp(max=1), :470
    const float w_old = w;
    const float w_new = __fadd_rn(w_old, fr.obs);
    tsdf = __fdiv_rn(__fadd_rn(__fmul_rn(w_old, tsdf), __fmul_rn(fr.obs, dist)), w_new);  // :479
    w = w_new;
    dirty = true;
  }
}

template <int SEM, bool COLOR>
__global__ void __launch_bounds__(kTsdfThreads) tsdf_integrate_kernel(const TsdfParams p) {
  __shared__ int s_list[kMaxFramesPerLaunch];
  __shared__ float s_zmax[kMaxFramesPerLaunch];
  __shared__ int s_wcount[kTsdfThreads / 32];
  __shared__ int s_box[6];
  __shared__ int s_n;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // ---- union of the frames' frustum boxes, in tile units -------------------------------------
  if (tid < 3) s_box[tid] = 0x7fffffff;
  else if (tid < 6) s_box[tid] = -1;
  __syncthreads();
  for (int f = tid; f < p.F; f += kTsdfThreads) {
    int lo[3], hi[3];
    const float zm = frame_zmax(p, f);
    s_zmax[f] = zm;
    if (frame_tile_box(p, p.frames[f], zm, lo, hi)) {
#pragma unroll
      for (int a = 0; a < 3; ++a) { atomicMin(&s_box[a], lo[a]); atomicMax(&s_box[3 + a], hi[a]); }
    }
  }
  __syncthreads();
  const int bx0 = s_box[0], by0 = s_box[1], bz0 = s_box[2];
  const int nbx = s_box[3] - bx0 + 1, nby = s_box[4] - by0 + 1, nbz = s_box[5] - bz0 + 1;
  if (nbx <= 0 || nby <= 0 || nbz <= 0) return;
  const int64_t ntiles = (int64_t)nbx * nby * nbz;

  // warp = 4x4x8 sub-box (x: 4 planes of one x-half, y-half, z-half); 8 consecutive z per row -> whole 32-byte sectors
  const int lz = (tid & 7) + 8 * (warp & 1), ly = ((tid >> 3) & 3) + 4 * ((warp >> 1) & 1), lxg = warp >> 2;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int tz = bz0 + (int)(t % nbz), ty = by0 + (int)((t / nbz) % nby), tx = bx0 + (int)(t / ((int64_t)nbz * nby));
    // tile box over voxel CENTRES, world space
    const int x0 = tx * kTileX, y0 = ty * kTileY, z0 = tz * kTileZ;
    const int x1 = min(p.dx, x0 + kTileX) - 1, y1 = min(p.dy, y0 + kTileY) - 1, z1 = min(p.dz, z0 + kTileZ) - 1;
    const float c[3] = {p.ox + (0.5f * (x0 + x1) + (float)p.xoff) * p.vs, p.oy + 0.5f * (y0 + y1) * p.vs,
                        p.oz + 0.5f * (z0 + z1) * p.vs};
    const float h[3] = {0.5f * (x1 - x0) * p.vs, 0.5f * (y1 - y0) * p.vs, 0.5f * (z1 - z0) * p.vs};
    // ---- ordered list of frames that can touch this tile --------------------------------------
    if (tid == 0) s_n = 0;
    __syncthreads();
    for (int f0 = 0; f0 < p.F; f0 += kTsdfThreads) {
      const int f = f0 + tid;
      bool keep = false;
      if (f < p.F) keep = tile_hits_frame(p.frames[f], s_zmax[f], p.trunc, c, h);
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) s_wcount[warp] = __popc(m);
      __syncthreads();
      int off = s_n;
      for (int w = 0; w < warp; ++w) off += s_wcount[w];
      if (keep) s_list[off + __popc(m & ((1u << lane) - 1u))] = f;
      __syncthreads();
      if (tid == 0) {
        int tot = 0;
        for (int w = 0; w < kTsdfThreads / 32; ++w) tot += s_wcount[w];
        s_n += tot;
      }
      __syncthreads();
    }
    const int nlist = s_n;
    if (nlist == 0) continue;
    // ---- tile-resident voxels --------------------------------------------------------------------
    const int z = z0 + lz, y = y0 + ly;
    const bool rowok = (z < p.dz) && (y < p.dy);
    float tv[kVoxPerThread], wv[kVoxPerThread], cv[kVoxPerThread];
    bool dirty[kVoxPerThread], inb[kVoxPerThread];
    int64_t idx[kVoxPerThread];
#pragma unroll
    for (int i = 0; i < kVoxPerThread; ++i) {
      const int x = x0 + lxg * kVoxPerThread + i;
      inb[i] = rowok && (x < p.dx);
      idx[i] = ((int64_t)x * p.dy + y) * p.dz + z;
      dirty[i] = false;
      tv[i] = 1.0f; wv[i] = 0.0f; cv[i] = 0.0f;
      if (inb[i]) {
        tv[i] = p.tsdf[idx[i]];
        wv[i] = p.weight[idx[i]];
        if (COLOR) cv[i] = p.color[idx[i]];
      }
    }
    // this warp's sub-box (voxel centres), for the second-level cull
    const int sx0 = x0 + 4 * lxg, sy0 = y0 + 4 * ((warp >> 1) & 1), sz0 = z0 + 8 * (warp & 1);
    const float wc[3] = {p.ox + ((float)(sx0 + p.xoff) + 1.5f) * p.vs, p.oy + ((float)sy0 + 1.5f) * p.vs,
                         p.oz + ((float)sz0 + 3.5f) * p.vs};
    const float wh[3] = {1.5f * p.vs, 1.5f * p.vs, 3.5f * p.vs};
    for (int li = 0; li < nlist; ++li) {
      const int f = s_list[li];
      const Frame& fr = p.frames[f];
      const float zm = s_zmax[f];
      if ((p.variant & 1) && !tile_hits_frame(fr, zm, p.trunc, wc, wh)) continue;  // warp-uniform
      const float zfar = (zm + p.trunc) * 1.000001f + 1e-6f;
      const float* depth = p.depth + (int64_t)f * p.H * p.W;
      const float* cimg = COLOR ? p.cimg + (int64_t)f * p.H * p.W : nullptr;
#pragma unroll
      for (int i = 0; i < kVoxPerThread; ++i) {
        if (inb[i])
          integrate_voxel<SEM, COLOR>(p, fr, depth, cimg, zfar, (float)(p.xoff + x0 + lxg * kVoxPerThread + i), (float)y,
                                      (float)z, tv[i], wv[i], cv[i], dirty[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < kVoxPerThread; ++i) {
      if (dirty[i]) {
        p.tsdf[idx[i]] = tv[i];
        p.weight[idx[i]] = wv[i];
        if (COLOR) p.color[idx[i]] = cv[i];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Host-side staging copy.  TSDFVolume.integrate() receives a pageable numpy frame (1.2 MB at 480x640); measured on the
// B200 box a single-threaded memcpy into the pinned ring takes ~130 us and is 90 % of the per-frame cost of the
// reference-shaped API (profiles/r01c_tsdf_host_profile.txt).  A few persistent helper threads split the copy.
// ------------------------------------------------------------------------------------------------
class CopyPool {
 public:
  static CopyPool& get() {
    static CopyPool* pool = new CopyPool();  // intentionally leaked: no destructor order problems at exit
    return *pool;
  }
  void copy(void* dst, const void* src, size_t bytes) {
    const int nw = (int)workers_.size();
    if (nw == 0 || bytes < (256u << 10)) { memcpy(dst, src, bytes); return; }
    std::lock_guard<std::mutex> one_caller(call_mu_);  // the job slot below is single-entry
    const size_t part = ((bytes / (size_t)(nw + 1)) + 4095) & ~(size_t)4095;
    {
      std::lock_guard<std::mutex> lk(mu_);
      dst_ = static_cast<char*>(dst); src_ = static_cast<const char*>(src); bytes_ = bytes; part_ = part;
      pending_ = nw;
      ++generation_;
    }
    cv_.notify_all();
    memcpy(dst, src, part < bytes ? part : bytes);  // the caller takes slice 0
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [&] { return pending_ == 0; });
  }

 private:
  CopyPool() {
    int n = 3;
    if (const char* e = getenv("D3M_COPY_THREADS")) n = atoi(e);
    const unsigned hc = std::thread::hardware_concurrency();
    if (hc > 0 && (unsigned)n + 1 > hc) n = (int)hc - 1;
    if (n < 0) n = 0;
    if (n > 15) n = 15;
    for (int i = 0; i < n; ++i) workers_.emplace_back([this, i] { run(i + 1); });
    for (auto& t : workers_) t.detach();
  }
  void run(int slice) {
    unsigned long long seen = 0;
    for (;;) {
      char* d; const char* s; size_t bytes, part;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return generation_ != seen; });
        seen = generation_;
        d = dst_; s = src_; bytes = bytes_; part = part_;
      }
      const size_t off = part * (size_t)slice;
      if (off < bytes) memcpy(d + off, s + off, (off + part <= bytes) ? part : bytes - off);
      {
        std::lock_guard<std::mutex> lk(mu_);
        --pending_;
      }
      done_.notify_one();
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_, call_mu_;
  std::condition_variable cv_, done_;
  char* dst_ = nullptr;
  const char* src_ = nullptr;
  size_t bytes_ = 0, part_ = 0;
  int pending_ = 0;
  unsigned long long generation_ = 0;
};

__global__ void fill_kernel(float* p, float v, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace d3m

using namespace d3m;

struct d3m_tsdf {
  int dx, dy, dz, xoff, device, sms;
  float origin[3], vs, trunc;
  float *tsdf, *weight, *color;
  size_t nvox;
  // pinned staging ring for host frames (depth + colour) and their device landing buffers
  float* pinned[kRing];
  float* dframe[kRing];
  cudaEvent_t ev[kRing];
  size_t frame_cap;  // floats per slot (2*H*W)
  int ring_pos;
  // frame tables
  Frame* h_frames[kRing];
  Frame* d_frames;
  float* d_partial;
  int frames_cap;
  cudaEvent_t fev[kRing];
  int fpos;
  int last_launches;
};

static void mat3_mul_vec(const double A[9], const double v[3], double o[3]) {
  for (int r = 0; r < 3; ++r) o[r] = A[3 * r] * v[0] + A[3 * r + 1] * v[1] + A[3 * r + 2] * v[2];
}

// Build the device-side description of one frame.  `pose16`: cam->world (kernel semantics) or world->cam
// (torch semantics).  Cull geometry is computed in double and is deliberately conservative.
static int build_frame(Frame& fr, const float* K9, const float* pose16, float obs, int H, int W, int sem) {
  fr.fx = K9[0]; fr.cx = K9[2]; fr.fy = K9[4]; fr.cy = K9[5];
  for (int i = 0; i < 12; ++i) fr.T[i] = pose16[i];
  fr.obs = obs;
  D3M_REQUIRE(fr.fx > 0.f && fr.fy > 0.f, D3M_ERR_ARG, "tsdf: focal lengths must be positive");
  // world->cam: cam = A p + a ; cam->world: p = Ai cam + centre
  double A[9], a[3], Ai[9], ctr[3];
  if (sem == D3M_TSDF_KERNEL_SEMANTICS) {
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) { Ai[3 * r + c] = pose16[4 * r + c]; A[3 * c + r] = pose16[4 * r + c]; }
    for (int r = 0; r < 3; ++r) ctr[r] = pose16[4 * r + 3];
    for (int r = 0; r < 3; ++r) a[r] = -(A[3 * r] * ctr[0] + A[3 * r + 1] * ctr[1] + A[3 * r + 2] * ctr[2]);
  } else {
    for (int r = 0; r < 3; ++r) {
      for (int c = 0; c < 3; ++c) A[3 * r + c] = pose16[4 * r + c];
      a[r] = pose16[4 * r + 3];
    }
    // general 3x3 inverse (the reference only assumes an invertible matrix)
    const double det = A[0] * (A[4] * A[8] - A[5] * A[7]) - A[1] * (A[3] * A[8] - A[5] * A[6]) +
                       A[2] * (A[3] * A[7] - A[4] * A[6]);
    D3M_REQUIRE(fabs(det) > 1e-12, D3M_ERR_ARG, "tsdf: singular world->camera matrix");
    Ai[0] = (A[4] * A[8] - A[5] * A[7]) / det; Ai[1] = (A[2] * A[7] - A[1] * A[8]) / det; Ai[2] = (A[1] * A[5] - A[2] * A[4]) / det;
    Ai[3] = (A[5] * A[6] - A[3] * A[8]) / det; Ai[4] = (A[0] * A[8] - A[2] * A[6]) / det; Ai[5] = (A[2] * A[3] - A[0] * A[5]) / det;
    Ai[6] = (A[3] * A[7] - A[4] * A[6]) / det; Ai[7] = (A[1] * A[6] - A[0] * A[7]) / det; Ai[8] = (A[0] * A[4] - A[1] * A[3]) / det;
    double na[3] = {-a[0], -a[1], -a[2]};
    mat3_mul_vec(Ai, na, ctr);
  }
  for (int r = 0; r < 3; ++r) fr.centre[r] = (float)ctr[r];
  const double us[2] = {-0.5 - 0.01, W - 0.5 + 0.01}, vs_[2] = {-0.5 - 0.01, H - 0.5 + 0.01};
  for (int k = 0; k < 4; ++k) {
    const double ray[3] = {(us[k & 1] - fr.cx) / fr.fx, (vs_[k >> 1] - fr.cy) / fr.fy, 1.0};
    double d[3];
    mat3_mul_vec(Ai, ray, d);
    for (int r = 0; r < 3; ++r) fr.dirs[k][r] = (float)d[r];
  }
  // camera-space half-spaces n.cam + d >= 0
  const double pc[5][3] = {{0, 0, 1},
                           {fr.fx, 0, fr.cx + 0.5},
                           {-fr.fx, 0, W - 0.5 - fr.cx},
                           {0, fr.fy, fr.cy + 0.5},
                           {0, -fr.fy, H - 0.5 - fr.cy}};
  for (int k = 0; k < 5; ++k) {
    double nw[3];
    for (int c = 0; c < 3; ++c) nw[c] = A[c] * pc[k][0] + A[3 + c] * pc[k][1] + A[6 + c] * pc[k][2];  // A^T n
    double dw = pc[k][0] * a[0] + pc[k][1] * a[1] + pc[k][2] * a[2];
    const double len = sqrt(nw[0] * nw[0] + nw[1] * nw[1] + nw[2] * nw[2]);
    D3M_REQUIRE(len > 0, D3M_ERR_ARG, "tsdf: degenerate camera");
    for (int c = 0; c < 3; ++c) fr.planes[k][c] = (float)(nw[c] / len);
    fr.planes[k][3] = (float)(dw / len);
  }
  {
    // far: cam_z <= zmax  <=>  -(A^T e_z).p - a_z + zmax >= 0 ; |A^T e_z| may differ from 1 for non-rigid input
    double nw[3] = {-A[6], -A[7], -A[8]};
    const double len = sqrt(nw[0] * nw[0] + nw[1] * nw[1] + nw[2] * nw[2]);
    const double s = len < 1.0 ? 1.0 : len;  // keep the test conservative: never shrink zmax
    for (int c = 0; c < 3; ++c) fr.far_n[c] = (float)(nw[c] / s);
    fr.far_d0 = (float)(-a[2] / s);
  }
  return D3M_OK;
}

static int tsdf_launch(d3m_tsdf* h, const float* depth, const float* cimg, int F, int H, int W, const Frame* d_frames,
                       int flags, cudaStream_t stream) {
  TsdfParams p;
  p.tsdf = h->tsdf; p.weight = h->weight; p.color = h->color;
  p.dx = h->dx; p.dy = h->dy; p.dz = h->dz; p.xoff = h->xoff;
  {
    static const int variant = getenv("D3M_TSDF_VARIANT") ? atoi(getenv("D3M_TSDF_VARIANT")) : 0;
    p.variant = variant;
  }
  p.ox = h->origin[0]; p.oy = h->origin[1]; p.oz = h->origin[2];
  p.vs = h->vs; p.trunc = h->trunc;
  p.frames = d_frames; p.F = F; p.depth = depth; p.cimg = cimg; p.H = H; p.W = W;
  p.partial_max = h->d_partial;
  {
    LaunchScope ls("tsdf_prep", stream);
    tsdf_prep_kernel<<<dim3(kPrepParts, F), 256, 0, stream>>>(depth, H * W, h->d_partial);
  }
  D3M_CUDA_CHECK(cudaGetLastError());
  const int sem = flags & 1;
  const bool color = (flags & D3M_TSDF_WITH_COLOR) && cimg != nullptr && sem == D3M_TSDF_KERNEL_SEMANTICS;
  const int grid = h->sms * 4;
  LaunchScope ls("tsdf_integrate", stream);
  if (sem == D3M_TSDF_KERNEL_SEMANTICS) {
    if (color) tsdf_integrate_kernel<D3M_TSDF_KERNEL_SEMANTICS, true><<<grid, kTsdfThreads, 0, stream>>>(p);
    else tsdf_integrate_kernel<D3M_TSDF_KERNEL_SEMANTICS, false><<<grid, kTsdfThreads, 0, stream>>>(p);
  } else {
    tsdf_integrate_kernel<D3M_TSDF_TORCH_SEMANTICS, false><<<grid, kTsdfThreads, 0, stream>>>(p);
  }
  D3M_CUDA_CHECK(cudaGetLastError());
  h->last_launches += 2;
  return D3M_OK;
}

static int ensure_frames(d3m_tsdf* h, int F) {
  if (F <= h->frames_cap) return D3M_OK;
  int cap = h->frames_cap ? h->frames_cap : 16;
  while (cap < F) cap *= 2;
  if (cap > kMaxFramesPerLaunch) cap = kMaxFramesPerLaunch;
  D3M_CUDA_CHECK(cudaDeviceSynchronize());
  for (int i = 0; i < kRing; ++i) {
    if (h->h_frames[i]) cudaFreeHost(h->h_frames[i]);
    D3M_CUDA_CHECK(cudaMallocHost(&h->h_frames[i], sizeof(Frame) * cap));
  }
  if (h->d_frames) cudaFree(h->d_frames);
  if (h->d_partial) cudaFree(h->d_partial);
  D3M_CUDA_CHECK(cudaMalloc(&h->d_frames, sizeof(Frame) * cap * kRing));
  D3M_CUDA_CHECK(cudaMalloc(&h->d_partial, sizeof(float) * cap * kPrepParts));
  h->frames_cap = cap;
  return D3M_OK;
}

extern "C" int d3m_tsdf_create(int dim_x, int dim_y, int dim_z, const float* origin3_host, float voxel_size,
                               float trunc_margin, int device, d3m_tsdf** out_handle) {
  return d3m_tsdf_create_slab(dim_x, dim_y, dim_z, 0, origin3_host, voxel_size, trunc_margin, device, out_handle);
}

extern "C" int d3m_tsdf_create_slab(int dim_x, int dim_y, int dim_z, int x_begin, const float* origin3_host,
                                    float voxel_size, float trunc_margin, int device, d3m_tsdf** out_handle) {
  D3M_REQUIRE(out_handle, D3M_ERR_ARG, "tsdf_create: NULL out_handle");
  *out_handle = nullptr;
  D3M_REQUIRE(d3m_device_count() > 0, D3M_ERR_NO_DEVICE, "tsdf_create: no CUDA device (there is no CPU fallback)");
  D3M_REQUIRE(dim_x > 0 && dim_y > 0 && dim_z > 0 && x_begin >= 0 && origin3_host && voxel_size > 0.f &&
                  trunc_margin > 0.f,
              D3M_ERR_ARG, "tsdf_create: bad arguments");
  DeviceGuard dev_guard(device);
  D3M_CUDA_CHECK(dev_guard.error());
  d3m_tsdf* h = new (std::nothrow) d3m_tsdf();
  D3M_REQUIRE(h, D3M_ERR_ARG, "tsdf_create: out of host memory");
  memset(h, 0, sizeof(*h));
  h->dx = dim_x; h->dy = dim_y; h->dz = dim_z; h->xoff = x_begin; h->device = device;
  h->vs = voxel_size; h->trunc = trunc_margin;
  memcpy(h->origin, origin3_host, 12);
  h->nvox = (size_t)dim_x * dim_y * dim_z;
  cudaDeviceGetAttribute(&h->sms, cudaDevAttrMultiProcessorCount, device);
  cudaError_t e = cudaMalloc(&h->tsdf, h->nvox * 4);
  if (e == cudaSuccess) e = cudaMalloc(&h->weight, h->nvox * 4);
  if (e == cudaSuccess) e = cudaMalloc(&h->color, h->nvox * 4);
  for (int i = 0; i < kRing && e == cudaSuccess; ++i) {
    e = cudaEventCreateWithFlags(&h->ev[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->fev[i], cudaEventDisableTiming);
  }
  if (e != cudaSuccess) { d3m_tsdf_destroy(h); return cuda_fail(e, "tsdf_create allocations"); }
  int rc = d3m_tsdf_reset(h, nullptr);
  if (rc == D3M_OK) rc = ensure_frames(h, 16);
  if (rc != D3M_OK) { d3m_tsdf_destroy(h); return rc; }
  cudaDeviceSynchronize();
  *out_handle = h;
  return D3M_OK;
}

extern "C" int d3m_tsdf_destroy(d3m_tsdf* h) {
  if (!h) return D3M_OK;
  DeviceGuard dev_guard(h->device);
  cudaDeviceSynchronize();
  cudaFree(h->tsdf); cudaFree(h->weight); cudaFree(h->color);
  for (int i = 0; i < kRing; ++i) {
    if (h->pinned[i]) cudaFreeHost(h->pinned[i]);
    if (h->dframe[i]) cudaFree(h->dframe[i]);
    if (h->h_frames[i]) cudaFreeHost(h->h_frames[i]);
    if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    if (h->fev[i]) cudaEventDestroy(h->fev[i]);
  }
  cudaFree(h->d_frames); cudaFree(h->d_partial);
  delete h;
  return D3M_OK;
}

extern "C" int d3m_tsdf_reset(d3m_tsdf* h, void* stream_) {
  D3M_REQUIRE(h, D3M_ERR_ARG, "tsdf_reset: NULL handle");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DeviceGuard dev_guard(h->device);
  D3M_CUDA_CHECK(dev_guard.error());
  {
    LaunchScope ls("tsdf_fill", stream);
    fill_kernel<<<h->sms * 8, 256, 0, stream>>>(h->tsdf, 1.0f, (int64_t)h->nvox);
  }  // tsdf_volume.py:50
  D3M_CUDA_CHECK(cudaGetLastError());
  D3M_CUDA_CHECK(cudaMemsetAsync(h->weight, 0, h->nvox * 4, stream));
  D3M_CUDA_CHECK(cudaMemsetAsync(h->color, 0, h->nvox * 4, stream));
  return D3M_OK;
}

extern "C" int d3m_tsdf_rebase(d3m_tsdf* h, const float* origin3_host, float voxel_size, float trunc_margin,
                               void* stream) {
  D3M_REQUIRE(h && origin3_host && voxel_size > 0.f && trunc_margin > 0.f, D3M_ERR_ARG, "tsdf_rebase: bad arguments");
  memcpy(h->origin, origin3_host, 12);
  h->vs = voxel_size;
  h->trunc = trunc_margin;
  return d3m_tsdf_reset(h, stream);
}

static int upload_frames(d3m_tsdf* h, int F, int H, int W, const float* intr9_host, int intr_per_frame,
                         const float* pose16_host, const float* obs_host, float obs_scalar, int flags,
                         cudaStream_t stream, const Frame** d_out) {
  int rc = ensure_frames(h, F);
  if (rc != D3M_OK) return rc;
  const int slot = h->fpos;
  h->fpos = (h->fpos + 1) % kRing;
  D3M_CUDA_CHECK(cudaEventSynchronize(h->fev[slot]));  // previous copy out of this pinned table finished
  Frame* hf = h->h_frames[slot];
  for (int f = 0; f < F; ++f) {
    rc = build_frame(hf[f], intr9_host + (intr_per_frame ? 9 * f : 0), pose16_host + 16 * f,
                     obs_host ? obs_host[f] : obs_scalar, H, W, flags & 1);
    if (rc != D3M_OK) return rc;
  }
  Frame* d = h->d_frames + (size_t)slot * h->frames_cap;
  D3M_CUDA_CHECK(cudaMemcpyAsync(d, hf, sizeof(Frame) * F, cudaMemcpyHostToDevice, stream));
  D3M_CUDA_CHECK(cudaEventRecord(h->fev[slot], stream));
  *d_out = d;
  return D3M_OK;
}

extern "C" int d3m_tsdf_integrate_device(d3m_tsdf* h, const float* depth, const float* color, int n_frames, int H,
                                         int W, const float* intr9_host, int intr_per_frame,
                                         const float* pose16_host, const float* obs_weight_host, int flags,
                                         void* stream_) {
  D3M_REQUIRE(h && depth && intr9_host && pose16_host && n_frames >= 0 && H > 0 && W > 0, D3M_ERR_ARG,
              "tsdf_integrate: bad arguments");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DeviceGuard dev_guard(h->device);
  D3M_CUDA_CHECK(dev_guard.error());
  h->last_launches = 0;
  for (int f0 = 0; f0 < n_frames; f0 += kMaxFramesPerLaunch) {
    const int F = (n_frames - f0) < kMaxFramesPerLaunch ? (n_frames - f0) : kMaxFramesPerLaunch;
    const Frame* d_frames = nullptr;
    int rc = upload_frames(h, F, H, W, intr9_host + (intr_per_frame ? 9 * f0 : 0), intr_per_frame,
                           pose16_host + 16 * f0, obs_weight_host ? obs_weight_host + f0 : nullptr, 1.0f, flags, stream,
                           &d_frames);
    if (rc != D3M_OK) return rc;
    rc = tsdf_launch(h, depth + (size_t)f0 * H * W, color ? color + (size_t)f0 * H * W : nullptr, F, H, W, d_frames,
                     flags, stream);
    if (rc != D3M_OK) return rc;
  }
  return D3M_OK;
}

extern "C" int d3m_tsdf_integrate_host(d3m_tsdf* h, const float* depth_host, const float* color_host, int H, int W,
                                       const float* intr9_host, const float* pose16_host, float obs_weight,
                                       int flags, void* stream_) {
  D3M_REQUIRE(h && depth_host && intr9_host && pose16_host && H > 0 && W > 0, D3M_ERR_ARG,
              "tsdf_integrate_host: bad arguments");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DeviceGuard dev_guard(h->device);
  D3M_CUDA_CHECK(dev_guard.error());
  const size_t hw = (size_t)H * W;
  if (h->frame_cap < 2 * hw) {
    D3M_CUDA_CHECK(cudaDeviceSynchronize());
    for (int i = 0; i < kRing; ++i) {
      if (h->pinned[i]) cudaFreeHost(h->pinned[i]);
      if (h->dframe[i]) cudaFree(h->dframe[i]);
      h->pinned[i] = nullptr; h->dframe[i] = nullptr;
      D3M_CUDA_CHECK(cudaMallocHost(&h->pinned[i], 2 * hw * 4));
      D3M_CUDA_CHECK(cudaMalloc(&h->dframe[i], 2 * hw * 4));
    }
    h->frame_cap = 2 * hw;
  }
  const int slot = h->ring_pos;
  h->ring_pos = (h->ring_pos + 1) % kRing;
  D3M_CUDA_CHECK(cudaEventSynchronize(h->ev[slot]));
  const bool with_color = color_host != nullptr && (flags & D3M_TSDF_WITH_COLOR);
  CopyPool::get().copy(h->pinned[slot], depth_host, hw * 4);
  if (with_color) CopyPool::get().copy(h->pinned[slot] + hw, color_host, hw * 4);
  D3M_CUDA_CHECK(cudaMemcpyAsync(h->dframe[slot], h->pinned[slot], (with_color ? 2 : 1) * hw * 4,
                                 cudaMemcpyHostToDevice, stream));
  D3M_CUDA_CHECK(cudaEventRecord(h->ev[slot], stream));
  h->last_launches = 0;
  const Frame* d_frames = nullptr;
  int rc = upload_frames(h, 1, H, W, intr9_host, 0, pose16_host, nullptr, obs_weight, flags, stream, &d_frames);
  if (rc != D3M_OK) return rc;
  return tsdf_launch(h, h->dframe[slot], with_color ? h->dframe[slot] + hw : nullptr, 1, H, W, d_frames, flags, stream);
}

extern "C" int d3m_tsdf_volumes(d3m_tsdf* h, float** tsdf, float** weight, float** color) {
  D3M_REQUIRE(h, D3M_ERR_ARG, "tsdf_volumes: NULL handle");
  if (tsdf) *tsdf = h->tsdf;
  if (weight) *weight = h->weight;
  if (color) *color = h->color;
  return D3M_OK;
}

// ---- d3m_upload: pageable host memory -> device through a pinned two-slot ring, staged by the copy pool ----------
namespace {
struct UploadRing {
  static constexpr size_t kChunk = 8u << 20;
  unsigned char* buf[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};
  bool used[2] = {false, false};
  int next = 0;  // slots alternate across calls too, so that back-to-back small uploads overlap staging and DMA
};
std::mutex g_upload_mu;
std::map<int, UploadRing> g_upload_rings;  // per device
}  // namespace

extern "C" int d3m_upload(const void* host_src, void* dev_dst, size_t bytes, void* stream_) {
  D3M_REQUIRE(d3m_device_count() > 0, D3M_ERR_NO_DEVICE, "d3m_upload: no CUDA device");
  if (bytes == 0) return D3M_OK;
  D3M_REQUIRE(host_src && dev_dst, D3M_ERR_ARG, "d3m_upload: NULL pointer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  int dev = 0;
  D3M_CUDA_CHECK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_upload_mu);
  UploadRing& r = g_upload_rings[dev];
  for (int i = 0; i < 2; ++i) {  // lazily, and per slot: a failed allocation leaves the ring consistent for the next call
    if (!r.buf[i]) D3M_CUDA_CHECK(cudaHostAlloc(reinterpret_cast<void**>(&r.buf[i]), UploadRing::kChunk, cudaHostAllocDefault));
    if (!r.done[i]) D3M_CUDA_CHECK(cudaEventCreateWithFlags(&r.done[i], cudaEventDisableTiming));
  }
  const unsigned char* src = static_cast<const unsigned char*>(host_src);
  unsigned char* dst = static_cast<unsigned char*>(dev_dst);
  for (size_t off = 0; off < bytes; off += UploadRing::kChunk) {
    const int slot = r.next;
    r.next ^= 1;
    const size_t n = bytes - off < UploadRing::kChunk ? bytes - off : UploadRing::kChunk;
    if (r.used[slot]) D3M_CUDA_CHECK(cudaEventSynchronize(r.done[slot]));  // the DMA that last read this slot
    CopyPool::get().copy(r.buf[slot], src + off, n);
    D3M_CUDA_CHECK(cudaMemcpyAsync(dst + off, r.buf[slot], n, cudaMemcpyHostToDevice, stream));
    D3M_CUDA_CHECK(cudaEventRecord(r.done[slot], stream));
    r.used[slot] = true;
  }
  return D3M_OK;
}

extern "C" int d3m_tsdf_download(d3m_tsdf* h, float* tsdf_host, float* weight_host, float* color_host, void* stream_) {
  D3M_REQUIRE(h, D3M_ERR_ARG, "tsdf_download: NULL handle");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  DeviceGuard dev_guard(h->device);
  D3M_CUDA_CHECK(dev_guard.error());
  if (tsdf_host) D3M_CUDA_CHECK(cudaMemcpyAsync(tsdf_host, h->tsdf, h->nvox * 4, cudaMemcpyDeviceToHost, stream));
  if (weight_host) D3M_CUDA_CHECK(cudaMemcpyAsync(weight_host, h->weight, h->nvox * 4, cudaMemcpyDeviceToHost, stream));
  if (color_host) D3M_CUDA_CHECK(cudaMemcpyAsync(color_host, h->color, h->nvox * 4, cudaMemcpyDeviceToHost, stream));
  D3M_CUDA_CHECK(cudaStreamSynchronize(stream));
  return D3M_OK;
}

extern "C" int d3m_tsdf_device(d3m_tsdf* h) { return h ? h->device : -1; }

extern "C" int d3m_tsdf_last_launches(d3m_tsdf* h) { return h ? h->last_launches : 0; }
