// TSDF fusion for sm_100a.  Replaces the PyCUDA `integrate` kernel + host loop of the reference
// (deep3dmap/core/tsdf/tsdf_volume.py:68-126, 210-256) and the torch-CPU `integrate()` used by the
// dataloader (tsdf_volume.py:437-482).
//
// Design: the reference launches one thread per voxel of the WHOLE volume for every frame (134 M threads at
// 512^3, >99 % of which exit at the frustum test) and copies seven arrays both ways per call.  Here
// three launches handle up to 512 frames (two for up to 32):
//   * tsdf_prep     reads every depth byte once, at HBM speed: max depth per 32x32-pixel block and per strip of blocks
//                   (the coarse depth grid the culls test against);
//   * tsdf_cull     for every 8x8x16-voxel tile inside the union of the frames' frustum boxes: one bit per frame that can
//                   touch it (frustum planes + coarse depth, conservative), tiles listed by their first frame word;
//   * tsdf_integrate  persistent, barrier-free: warps pull 4x4x8-voxel sub-boxes from one queue, keep the sub-box's
//                   tsdf / weight in REGISTERS while they apply the surviving frames in order (the running average is
//                   order dependent), and write back only what changed.  Volume bytes move once per launch instead of
//                   once per frame; a lane row is 8 consecutive z = whole 32-byte sectors.
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
constexpr int kTsdfThreads = 256;         // 8 independent warps; a warp works on one 4x4x8-voxel sub-box at a time
constexpr int kVoxPerThread = 4;          // a lane owns 4 consecutive x at one (y, z)
constexpr int kMaxFramesPerLaunch = 512;  // hot frame data of one launch is staged in shared memory (80 B per frame)
constexpr int kPrepParts = 16;            // strips of the coarse depth grid per frame (rows of blocks)
constexpr int kBlkCols = 32;              // blocks per strip at most
constexpr int kFrameAux = kPrepParts + kPrepParts * kBlkCols;  // floats per frame: strip maxima, then block maxima
constexpr int kRing = 4;
constexpr float kSlack = 0.02f;           // metres; conservative margin of the cull tests
constexpr int kMaxWords = kMaxFramesPerLaunch / 32;
constexpr int kCounters = 1 + kMaxWords;

// Host-side description of one frame (build_frame); packed for the device by pack_frames():
//   hot  (kHotFloats per frame, frame-major): what the per-voxel arithmetic reads -- staged in shared memory
//   cull (kCullFields x F, field-major):      what the conservative box tests read -- lane j <-> frame j, coalesced
struct Frame {
  float fx, fy, cx, cy;
  float T[12];       // kernel semantics: rows 0..2 of cam->world pose; torch semantics: rows 0..2 of world->cam
  float obs;
  float centre[3];   // camera centre, world
  float dirs[4][3];  // world-space rays through the 4 image corners (per unit camera depth)
  float planes[5][4];  // near, left, right, top, bottom: unit normal (world) and offset; inside if n.p + d >= 0
  float far_n[3];
  float far_d0;        // inside if far_n.p + far_d0 + zmax >= 0
  float Wc[12];        // world->camera, rows 0..2 (cull geometry only: fp32 copy of the fp64 matrix)
};
constexpr int kHotFloats = 20;  // fx fy cx cy | T[12] | obs, 3 pad  (five 128-bit shared-memory loads)
enum { kCfFx = 0, kCfFy, kCfCx, kCfCy, kCfCentre = 4, kCfDirs = 7, kCfPlanes = 19, kCfFarN = 39, kCfFarD0 = 42, kCfWc = 43,
       kCullFields = 55 };

struct TsdfParams {
  float* tsdf;
  float* weight;
  float* color;
  int dx, dy, dz;
  int xoff;             // global x index of local plane 0 (slab sharding), 0 otherwise
  int variant;          // A/B switches (D3M_TSDF_VARIANT, default 0 = everything on): 1 = no per-warp frame masks,
                        // 2 = no coarse-depth test in the culls, 4 = lazy sub-box load, 8 = branchy per-voxel code, 16 = cull kernel
                        // also for launches of <= 32 frames
  float ox, oy, oz, vs, trunc;
  const float* hot;     // (F, kHotFloats)
  const float* cull;    // (kCullFields, F)
  int F;
  const float* depth;   // (F,H,W)
  const float* cimg;    // (F,H,W) folded colour or NULL
  int H, W;
  int bsl, nbx, nby;    // coarse depth grid: blocks of (1 << bsl)^2 pixels, nbx x nby blocks per frame
  float* aux;           // (F, kFrameAux): [0,16) strip maxima, then nby x nbx block maxima (max depth, +inf if NaN inside)
  float* zmax;          // (F) largest depth of each frame (written by the cull kernel)
  // per-launch cull results (tsdf_cull_kernel -> tsdf_integrate_kernel)
  unsigned int* masks;      // (tiles of the volume, words): bit j of word w of a tile = frame 32w + j may touch the tile
  int words;                // ceil(F / 32)
  int* word_lists;          // (words, tiles of the volume): list w = the tiles whose FIRST non-empty mask word is w
  unsigned int* counters;   // [0] work queue of the integrate kernel, [1 + w] length of list w; cleared by the prep kernel
  int64_t n_vol_tiles;
  int tnx, tny, tnz;        // tiles of the volume
  int direct;               // 1 (launches of up to 32 frames): no cull kernel -- the integrate kernel walks every sub-box
                            // of the frames' union box itself and tests the frames per sub-box only
};

// Coarse depth grid of every frame: max depth per block of (1 << bsl)^2 pixels, and per strip (row) of blocks.  Reads
// every depth byte once, at HBM speed: one warp per block, 128-bit loads, four 128-byte row segments per warp step, a
// shuffle tree per block; the CTA (one warp per block of the strip) folds the strip maximum through shared memory.
// The maximum is taken on the BIT PATTERNS as signed integers: non-negative floats order like their bits, every negative
// float is a negative integer (clamped away by the initial 0: conservative, a voxel such a pixel updates has
// cam_z < trunc), and a NaN compares above +inf -- it becomes +inf, which disables the far culls for its block and
// strip: a NaN depth passes the reference's `depth == 0` / `diff < -trunc` tests (tsdf_volume.py:114,119) and can
// update voxels at any distance.
// The same launch clears the per-launch counters (list lengths, work queue).
__global__ void __launch_bounds__(32 * kBlkCols) tsdf_prep_kernel(const float* __restrict__ depth, int H, int W, int bsl,
                                                                 int nbx, int nby, float* __restrict__ aux,
                                                                 unsigned int* counters, int n_counters) {
  __shared__ int s_blk[kBlkCols];
  const int f = blockIdx.y, by = blockIdx.x;
  const int lane = threadIdx.x & 31, bx = threadIdx.x >> 5;   // blockDim.x = 32 * nbx
  if (blockIdx.x == 0 && blockIdx.y == 0)
    for (int i = threadIdx.x; i < n_counters; i += blockDim.x) counters[i] = 0u;
  float* a = aux + (size_t)f * kFrameAux;
  if (by >= nby) {
    if (threadIdx.x == 0) a[by] = 0.0f;
    return;
  }
  const int bs = 1 << bsl;
  const int y0 = by * bs, y1 = min(H, y0 + bs);
  const int x0 = bx * bs, x1 = min(W, x0 + bs);
  const float* d = depth + (size_t)f * H * W;
  int m = 0;
  if ((W & 3) == 0 && bsl == 5) {
    // 32 x 32 block: lane <-> (row r of 4, 16-byte column c4 of 8); the 8 row steps are independent loads in flight
    const int r = lane >> 3, x = x0 + (lane & 7) * 4;
#pragma unroll
    for (int k0 = 0; k0 < 8; k0 += 4) {
      int4 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int y = y0 + r + 4 * (k0 + k);
        v[k] = (y < y1 && x < x1) ? __ldg(reinterpret_cast<const int4*>(d + (size_t)y * W + x)) : make_int4(0, 0, 0, 0);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) m = max(max(m, max(v[k].x, v[k].y)), max(v[k].z, v[k].w));
    }
  } else if ((W & 3) == 0) {
    const int r = lane >> 3, c4 = (lane & 7) * 4;   // 4 rows x 32 columns per warp step
    for (int y = y0 + r; y < y1; y += 4)
      for (int x = x0 + c4; x < x1; x += 32) {
        const int4 v = __ldg(reinterpret_cast<const int4*>(d + (size_t)y * W + x));   // x < W and W % 4 == 0
        m = max(max(m, max(v.x, v.y)), max(v.z, v.w));
      }
  } else {
    for (int y = y0; y < y1; ++y)
      for (int x = x0 + lane; x < x1; x += 32) m = max(m, __float_as_int(__ldg(d + (size_t)y * W + x)));
  }
  m = __reduce_max_sync(0xffffffffu, m);
  m = min(m, 0x7f800000);   // NaN -> +inf
  if (lane == 0) {
    a[kPrepParts + by * nbx + bx] = __int_as_float(m);
    s_blk[bx] = m;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int sm = 0;
    for (int k = 0; k < nbx; ++k) sm = max(sm, s_blk[k]);
    a[by] = __int_as_float(sm);
  }
}

__device__ __forceinline__ float frame_zmax(const TsdfParams& p, int f) {
  float m = 0.0f;
  const float* a = p.aux + (size_t)f * kFrameAux;
#pragma unroll
  for (int i = 0; i < kPrepParts; ++i) m = fmaxf(m, a[i]);
  return m;  // 0 -> frame has no valid depth
}

__device__ __forceinline__ float cf(const TsdfParams& p, int field, int f) { return __ldg(p.cull + (size_t)field * p.F + f); }

// frustum box of frame f in tile coordinates (inclusive), false when empty
__device__ __forceinline__ bool frame_tile_box(const TsdfParams& p, int f, float zmaxd, int lo[3], int hi[3]) {
  if (!(zmaxd > 0.0f)) return false;
  const float zm = zmaxd + p.trunc + kSlack;
  float mn[3], mx[3], ctr[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) { ctr[a] = cf(p, kCfCentre + a, f); mn[a] = ctr[a]; mx[a] = ctr[a]; }
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float q = ctr[a] + zm * cf(p, kCfDirs + 3 * k + a, f);
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

// Conservative test: can ANY voxel centre inside the axis-aligned box (centre c, half extents h, world space) be updated
// by frame f?  (1) the five frustum planes and the far plane at the frame's largest depth; (2) the coarse depth grid:
// the box is bounded in camera space (centre through the fp32 world->camera matrix, extents through its absolute values),
// projected to a pixel rectangle, and compared with the largest depth of the blocks under that rectangle -- a voxel
// behind every measurement it could see by more than the truncation distance is skipped by the reference
// (`depth - cam_z < -trunc`, tsdf_volume.py:119).  All margins (kSlack, 1.5 pixels, the AABB) err on the keeping side, so
// results are unchanged bit for bit.  Lane j of a warp tests frame j: the field-major table makes every load coalesced.
__device__ __forceinline__ bool box_hits_frame(const TsdfParams& p, int f, float zmaxd, const float c[3], const float h[3],
                                               bool depth_grid) {
  if (!(zmaxd > 0.0f)) return false;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const float n0 = cf(p, kCfPlanes + 4 * k, f), n1 = cf(p, kCfPlanes + 4 * k + 1, f), n2 = cf(p, kCfPlanes + 4 * k + 2, f);
    const float dist = n0 * c[0] + n1 * c[1] + n2 * c[2] + cf(p, kCfPlanes + 4 * k + 3, f);
    const float reach = fabsf(n0) * h[0] + fabsf(n1) * h[1] + fabsf(n2) * h[2];
    if (dist + reach < -kSlack) return false;
  }
  {
    const float n0 = cf(p, kCfFarN, f), n1 = cf(p, kCfFarN + 1, f), n2 = cf(p, kCfFarN + 2, f);
    const float dist = n0 * c[0] + n1 * c[1] + n2 * c[2] + cf(p, kCfFarD0, f) + zmaxd + p.trunc;
    const float reach = fabsf(n0) * h[0] + fabsf(n1) * h[1] + fabsf(n2) * h[2];
    if (dist + reach < -kSlack) return false;
  }
  if (!depth_grid) return true;
  float M[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) M[i] = cf(p, kCfWc + i, f);
  const float ccx = M[0] * c[0] + M[1] * c[1] + M[2] * c[2] + M[3];
  const float ccy = M[4] * c[0] + M[5] * c[1] + M[6] * c[2] + M[7];
  const float ccz = M[8] * c[0] + M[9] * c[1] + M[10] * c[2] + M[11];
  const float ex = fabsf(M[0]) * h[0] + fabsf(M[1]) * h[1] + fabsf(M[2]) * h[2] + kSlack;
  const float ey = fabsf(M[4]) * h[0] + fabsf(M[5]) * h[1] + fabsf(M[6]) * h[2] + kSlack;
  const float ez = fabsf(M[8]) * h[0] + fabsf(M[9]) * h[1] + fabsf(M[10]) * h[2] + kSlack;
  const float zn = ccz - ez, zf = ccz + ez;
  if (!(zn > 0.05f)) return true;   // touches the camera plane: the projection is unbounded, keep
  // u = fx * x / z + cx over x in [ccx-ex, ccx+ex], z in [zn, zf]: extremes at the corners
  const float fx = cf(p, kCfFx, f), fy = cf(p, kCfFy, f), pcx = cf(p, kCfCx, f), pcy = cf(p, kCfCy, f);
  const float xl = ccx - ex, xh = ccx + ex, yl = ccy - ey, yh = ccy + ey;
  const float inv_n = 1.0f / zn, inv_f = 1.0f / zf;
  const float ul = fx * (xl * (xl < 0.0f ? inv_n : inv_f)) + pcx - 1.5f;
  const float uh = fx * (xh * (xh < 0.0f ? inv_f : inv_n)) + pcx + 1.5f;
  const float vl = fy * (yl * (yl < 0.0f ? inv_n : inv_f)) + pcy - 1.5f;
  const float vh = fy * (yh * (yh < 0.0f ? inv_f : inv_n)) + pcy + 1.5f;
  if (!(uh >= 0.0f && vh >= 0.0f && ul <= (float)(p.W - 1) && vl <= (float)(p.H - 1)))
    return !(ul == ul && vl == vl && uh == uh && vh == vh);   // wholly outside the image -> no voxel can pass; NaN -> keep
  const int bx0 = max(0, (int)fmaxf(ul, 0.0f)) >> p.bsl, bx1 = min(p.W - 1, (int)fminf(uh, (float)(p.W - 1))) >> p.bsl;
  const int by0 = max(0, (int)fmaxf(vl, 0.0f)) >> p.bsl, by1 = min(p.H - 1, (int)fminf(vh, (float)(p.H - 1))) >> p.bsl;
  const float* aux = p.aux + (size_t)f * kFrameAux + kPrepParts;
  float md = 0.0f;
  for (int by = by0; by <= by1; ++by)
    for (int bx = bx0; bx <= bx1; ++bx) md = fmaxf(md, __ldg(aux + by * p.nbx + bx));
  return zn <= md + p.trunc + kSlack;
}

// PTX cvt.rzi.s32.f32 is what `(int)` compiles to: saturating, NaN -> 0 (same as in the reference kernel)
__device__ __forceinline__ int f2i_rz(float x) { return __float2int_rz(x); }

// (int)roundf(u) for -0.5 < u < 2^22 (or NaN): round-to-nearest-even, then the one case where half-away-from-zero
// differs (an exact tie that went down) is corrected.  NaN stays NaN and converts to 0, like roundf.
__device__ __forceinline__ int round_half_away_to_int(float u) {
  float r = rintf(u);
  if (__fsub_rn(r, u) == -0.5f) r = __fadd_rn(r, 1.0f);
  return f2i_rz(r);
}

// One frame applied to the kVoxPerThread voxels of a lane (consecutive x, same y and z).  Arithmetic = the reference's,
// rounding step for rounding step; only the ORDER of the rejection tests is changed so that the cheap ones come first
// (cam_z < 0 before the two divisions; the image-bounds test on the un-rounded pixel coordinate: (int)roundf(u) < 0 <=>
// u <= -0.5 and (int)roundf(u) >= W <=> u >= W - 0.5, NaN falls through to pixel 0 exactly as cvt.rzi makes it).
// `fh`: the frame's hot data in shared memory (fx fy cx cy | T[12] | obs).  Returns the mask of voxels whose pixel lies in
// the image, with the pixel offset and cam_z of each.
template <int SEM>
__device__ __forceinline__ unsigned project_voxels(const TsdfParams& p, const float* __restrict__ fh,
                                                   const float ptx[kVoxPerThread], float pty, float ptz, unsigned inb,
                                                   float camz[kVoxPerThread], int pix[kVoxPerThread]) {
  const float4 k4 = *reinterpret_cast<const float4*>(fh);        // fx fy cx cy
  const float4 t0 = *reinterpret_cast<const float4*>(fh + 4);    // T[0..3]
  const float4 t1 = *reinterpret_cast<const float4*>(fh + 8);    // T[4..7]
  const float4 t2 = *reinterpret_cast<const float4*>(fh + 12);   // T[8..11]
  const float fx = k4.x, fy = k4.y, pcx = k4.z, pcy = k4.w;
  unsigned ok = 0u;
  if (SEM == D3M_TSDF_KERNEL_SEMANTICS) {
    // tsdf_volume.py:94-106 with the contraction of the reference build
    const float ty = __fsub_rn(pty, t1.w), tz = __fsub_rn(ptz, t2.w);
    const float ay0 = __fmul_rn(ty, t1.x), ay1 = __fmul_rn(ty, t1.y), ay2 = __fmul_rn(ty, t1.z);
#pragma unroll
    for (int i = 0; i < kVoxPerThread; ++i) {
      pix[i] = 0; camz[i] = 0.0f;
      if (!((inb >> i) & 1u)) continue;
      const float tx = __fsub_rn(ptx[i], t0.w);
      const float cz = __fmaf_rn(tz, t2.z, __fmaf_rn(tx, t0.z, ay2));
      if (cz < 0.0f) continue;                                     // :110 (last clause)
      const float cxm = __fmaf_rn(tz, t2.x, __fmaf_rn(tx, t0.x, ay0));
      const float u = __fmaf_rn(fx, __fdiv_rn(cxm, cz), pcx);
      if (u <= -0.5f || u >= (float)p.W - 0.5f) continue;          // :110 px < 0 || px >= W
      const float cym = __fmaf_rn(tz, t2.y, __fmaf_rn(tx, t0.y, ay1));
      const float v = __fmaf_rn(fy, __fdiv_rn(cym, cz), pcy);
      if (v <= -0.5f || v >= (float)p.H - 0.5f) continue;          // :110 py < 0 || py >= H
      pix[i] = round_half_away_to_int(v) * p.W + round_half_away_to_int(u);
      camz[i] = cz;
      ok |= 1u << i;
    }
  } else {
    // tsdf_volume.py:523 (world_c), :451-459; ptx/pty/ptz were formed as origin + vs * index (mul, add)
#pragma unroll
    for (int i = 0; i < kVoxPerThread; ++i) {
      pix[i] = 0; camz[i] = 0.0f;
      if (!((inb >> i) & 1u)) continue;
      const float wx = ptx[i];
      const float cz = __fadd_rn(__fmaf_rn(t2.z, ptz, __fmaf_rn(t2.y, pty, __fmul_rn(t2.x, wx))), t2.w);
      if (!(cz > 0.0f)) continue;                                  // :462 (last clause)
      const float cxm = __fadd_rn(__fmaf_rn(t0.z, ptz, __fmaf_rn(t0.y, pty, __fmul_rn(t0.x, wx))), t0.w);
      const float cym = __fadd_rn(__fmaf_rn(t1.z, ptz, __fmaf_rn(t1.y, pty, __fmul_rn(t1.x, wx))), t1.w);
      const float rx = rintf(__fadd_rn(__fdiv_rn(__fmul_rn(cxm, fx), cz), pcx));
      const float ry = rintf(__fadd_rn(__fdiv_rn(__fmul_rn(cym, fy), cz), pcy));
      if (!((rx >= 0.0f) && (rx < (float)p.W) && (ry >= 0.0f) && (ry < (float)p.H))) continue;   // :462
      pix[i] = (int)ry * p.W + (int)rx;
      camz[i] = cz;
      ok |= 1u << i;
    }
  }
  return ok;
}

// Second half: the sampled depths decide which voxels the frame updates and with which clamped observation.
template <int SEM>
__device__ __forceinline__ unsigned depth_test_voxels(const TsdfParams& p, unsigned ok, const float d[kVoxPerThread],
                                                      const float camz[kVoxPerThread], float dist[kVoxPerThread]) {
  unsigned upd = 0u;
#pragma unroll
  for (int i = 0; i < kVoxPerThread; ++i) {
    if (!((ok >> i) & 1u)) continue;
    const float diff = __fsub_rn(d[i], camz[i]);
    if (SEM == D3M_TSDF_KERNEL_SEMANTICS) {
      if (d[i] == 0.0f) continue;            // :114
      if (diff < -p.trunc) continue;         // :119
      dist[i] = fminf(__fdiv_rn(diff, p.trunc), 1.0f);
    } else {
      if (!(d[i] > 0.0f && diff >= -p.trunc)) continue;  // :471
      float t = __fdiv_rn(diff, p.trunc);
      if (t > 1.0f) t = 1.0f;                             // clamp(max=1), :470
      dist[i] = t;
    }
    upd |= 1u << i;
  }
  return upd;
}

// Straight-line variants of the two halves (kernel semantics only): every step is computed for all kVoxPerThread voxels
// and the rejection tests only form the mask at the end.  Same values, same tests; the early `continue`s of the versions
// above made ptxas emit one convergence region per voxel and test, which serialises the four independent division chains
// (issue-active 57 %); here the chains interleave.  The price is arithmetic on voxels that fail (a division by cam_z = 0
// yields inf / NaN, which the tests reject exactly like the reference's int conversion does).
__device__ __forceinline__ unsigned project_voxels_sl(const TsdfParams& p, const float* __restrict__ fh,
                                                      const float ptx[kVoxPerThread], float pty, float ptz, unsigned inb,
                                                      float camz[kVoxPerThread], int pix[kVoxPerThread]) {
  const float4 k4 = *reinterpret_cast<const float4*>(fh);
  const float4 t0 = *reinterpret_cast<const float4*>(fh + 4);
  const float4 t1 = *reinterpret_cast<const float4*>(fh + 8);
  const float4 t2 = *reinterpret_cast<const float4*>(fh + 12);
  const float ty = __fsub_rn(pty, t1.w), tz = __fsub_rn(ptz, t2.w);
  const float ay0 = __fmul_rn(ty, t1.x), ay1 = __fmul_rn(ty, t1.y), ay2 = __fmul_rn(ty, t1.z);
  const float wlim = (float)p.W - 0.5f, hlim = (float)p.H - 0.5f;
  float u[kVoxPerThread], v[kVoxPerThread];
#pragma unroll
  for (int i = 0; i < kVoxPerThread; ++i) {
    const float tx = __fsub_rn(ptx[i], t0.w);
    camz[i] = __fmaf_rn(tz, t2.z, __fmaf_rn(tx, t0.z, ay2));
    const float cxm = __fmaf_rn(tz, t2.x, __fmaf_rn(tx, t0.x, ay0));
    const float cym = __fmaf_rn(tz, t2.y, __fmaf_rn(tx, t0.y, ay1));
    u[i] = __fmaf_rn(k4.x, __fdiv_rn(cxm, camz[i]), k4.z);
    v[i] = __fmaf_rn(k4.y, __fdiv_rn(cym, camz[i]), k4.w);
  }
  unsigned ok = 0u;
#pragma unroll
  for (int i = 0; i < kVoxPerThread; ++i) {
    const bool good = ((inb >> i) & 1u) && !(camz[i] < 0.0f) && !(u[i] <= -0.5f || u[i] >= wlim) &&
                      !(v[i] <= -0.5f || v[i] >= hlim);                     // tsdf_volume.py:110
    pix[i] = good ? round_half_away_to_int(v[i]) * p.W + round_half_away_to_int(u[i]) : 0;
    ok |= good ? (1u << i) : 0u;
  }
  return ok;
}

__device__ __forceinline__ unsigned depth_test_voxels_sl(const TsdfParams& p, unsigned ok, const float d[kVoxPerThread],
                                                         const float camz[kVoxPerThread], float dist[kVoxPerThread]) {
  unsigned upd = 0u;
#pragma unroll
  for (int i = 0; i < kVoxPerThread; ++i) {
    const float diff = __fsub_rn(d[i], camz[i]);
    dist[i] = fminf(__fdiv_rn(diff, p.trunc), 1.0f);
    const bool good = ((ok >> i) & 1u) && !(d[i] == 0.0f) && !(diff < -p.trunc);   // :114, :119
    upd |= good ? (1u << i) : 0u;
  }
  return upd;
}

// ---- cull: which frames can touch which tile ---------------------------------------------------------------------
// Every CTA derives the union of the frames' frustum boxes (in tile units); the warps then stride over the tiles of that
// box: lane j tests frame 32w + j against the tile for every mask word w (one ballot per word); a tile with any bit set
// is appended to the work list of its FIRST non-empty word.  No block-level synchronisation after the prologue.
constexpr int kCullThreads = 256;
__global__ void __launch_bounds__(kCullThreads) tsdf_cull_kernel(const TsdfParams p) {
  __shared__ float s_zmax[kMaxFramesPerLaunch];
  __shared__ int s_box[6];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool use_grid = !(p.variant & 2);
  if (tid < 3) s_box[tid] = 0x7fffffff;
  else if (tid < 6) s_box[tid] = -1;
  __syncthreads();
  for (int f = tid; f < p.F; f += kCullThreads) {
    int lo[3], hi[3];
    const float zm = frame_zmax(p, f);
    s_zmax[f] = zm;
    if (blockIdx.x == 0) p.zmax[f] = zm;
    if (frame_tile_box(p, f, zm, lo, hi)) {
#pragma unroll
      for (int a = 0; a < 3; ++a) { atomicMin(&s_box[a], lo[a]); atomicMax(&s_box[3 + a], hi[a]); }
    }
  }
  __syncthreads();
  const int bx0 = s_box[0], by0 = s_box[1], bz0 = s_box[2];
  const int nbx = s_box[3] - bx0 + 1, nby = s_box[4] - by0 + 1, nbz = s_box[5] - bz0 + 1;
  if (nbx <= 0 || nby <= 0 || nbz <= 0) return;
  const int64_t ntiles = (int64_t)nbx * nby * nbz;
  const int64_t nwarps = (int64_t)gridDim.x * (kCullThreads / 32);
  for (int64_t t = (int64_t)blockIdx.x * (kCullThreads / 32) + warp; t < ntiles; t += nwarps) {
    const int tz = bz0 + (int)(t % nbz), ty = by0 + (int)((t / nbz) % nby), tx = bx0 + (int)(t / ((int64_t)nbz * nby));
    // tile box over voxel CENTRES, world space
    const int x0 = tx * kTileX, y0 = ty * kTileY, z0 = tz * kTileZ;
    const int x1 = min(p.dx, x0 + kTileX) - 1, y1 = min(p.dy, y0 + kTileY) - 1, z1 = min(p.dz, z0 + kTileZ) - 1;
    const float c[3] = {p.ox + (0.5f * (x0 + x1) + (float)p.xoff) * p.vs, p.oy + 0.5f * (y0 + y1) * p.vs,
                        p.oz + 0.5f * (z0 + z1) * p.vs};
    const float h[3] = {0.5f * (x1 - x0) * p.vs, 0.5f * (y1 - y0) * p.vs, 0.5f * (z1 - z0) * p.vs};
    const int64_t tl = ((int64_t)tx * p.tny + ty) * p.tnz + tz;
    int first = -1;
    for (int w = 0; w < p.words; ++w) {
      const int f = 32 * w + lane;
      bool keep = false;
      if (f < p.F) keep = box_hits_frame(p, f, s_zmax[f], c, h, use_grid);
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) p.masks[tl * p.words + w] = m;
      if (m != 0u && first < 0) first = w;
    }
    if (first >= 0 && lane == 0)
      p.word_lists[(size_t)first * p.n_vol_tiles + atomicAdd(&p.counters[1 + first], 1u)] = (int)tl;
  }
}

// ---- integrate: warps are independent workers ----------------------------------------------------------------------
// Work item = one 4x4x8-voxel sub-box of a listed tile: a lane owns 4 consecutive x at one (y, z), 8 consecutive z per lane
// row = whole 32-byte sectors.  A warp pulls items from ONE queue, keeps the sub-box's tsdf / weight in registers while it
// walks the tile's frame mask word by word (lane j <-> bit j: refined for the sub-box, then the surviving frames applied in
// order), and writes back what changed -- volume bytes move once per launch.
// Queue order = the tiles' FIRST frame word: at any moment the GPU works on sub-boxes that start with the same ~32 frames
// and (a camera moves continuously) end two or three words later, so the depth images in use (~100-150 MB) mostly stay in
// the 126 MB L2.  Measured alternatives (profiles/r02_tsdf_variants.txt): tile order as listed re-read the 369 MB of depth
// 3.7 times (1.37 GB of DRAM traffic, 62 % of the HBM peak); strict frame-word-major items with a per-sub-box progress word
// got it down to 0.49 GB but lost the gain to waiting (a word touches only a sector of the scene, < 1.5 waves of items,
// so the predecessor of an item is usually still running).
// The only __syncthreads of the kernel is the one after staging the frames' hot data in shared memory: with CTA-wide
// frame lists a third of the warp stalls were barrier waits (profiles/r02e_tsdf_ncu.txt).
template <int SEM, bool COLOR, bool SL>
__global__ void __launch_bounds__(kTsdfThreads, 4) tsdf_integrate_kernel(const TsdfParams p) {
  extern __shared__ __align__(16) float s_hot[];   // (F, kHotFloats)
  __shared__ unsigned s_first[kMaxWords + 1];      // first queue item of every list
  __shared__ float s_zmax32[32];                   // direct mode: largest depth of each frame
  __shared__ int s_box[6];                         // direct mode: union of the frames' frustum boxes, in tiles
  for (int i = threadIdx.x; i < p.F * (kHotFloats / 4); i += kTsdfThreads)
    reinterpret_cast<float4*>(s_hot)[i] = __ldg(reinterpret_cast<const float4*>(p.hot) + i);
  if (p.direct) {
    if (threadIdx.x < 3) s_box[threadIdx.x] = 0x7fffffff;
    else if (threadIdx.x < 6) s_box[threadIdx.x] = -1;
    __syncthreads();
    if ((int)threadIdx.x < p.F) {
      int lo[3], hi[3];
      const float zm = frame_zmax(p, threadIdx.x);
      s_zmax32[threadIdx.x] = zm;
      if (frame_tile_box(p, threadIdx.x, zm, lo, hi)) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { atomicMin(&s_box[a], lo[a]); atomicMax(&s_box[3 + a], hi[a]); }
      }
    }
  } else if (threadIdx.x == 0) {
    unsigned acc = 0u;
    for (int w = 0; w < p.words; ++w) { s_first[w] = acc; acc += __ldcg(&p.counters[1 + w]) * 8u; }
    s_first[p.words] = acc;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const bool use_grid = !(p.variant & 2);
  const int dbx0 = s_box[0], dby0 = s_box[1], dbz0 = s_box[2];
  const int dnx = s_box[3] - dbx0 + 1, dny = s_box[4] - dby0 + 1, dnz = s_box[5] - dbz0 + 1;
  unsigned n_items;
  if (p.direct) n_items = (dnx > 0 && dny > 0 && dnz > 0) ? (unsigned)min((int64_t)dnx * dny * dnz * 8, (int64_t)0x7fffffff) : 0u;
  else n_items = s_first[p.words];
  int wl = 0;
  for (;;) {
    unsigned item = 0u;
    if (lane == 0) item = atomicAdd(&p.counters[0], 1u);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= n_items) break;
    int tx, ty, tz, sub;
    int64_t tl;
    if (p.direct) {
      const unsigned t = item >> 3;
      sub = (int)(item & 7u);
      tz = dbz0 + (int)(t % dnz); ty = dby0 + (int)((t / dnz) % dny); tx = dbx0 + (int)(t / ((unsigned)dnz * dny));
      tl = ((int64_t)tx * p.tny + ty) * p.tnz + tz;
    } else {
      while (item >= s_first[wl + 1]) ++wl;          // items only grow: the list index never goes back
      const unsigned local = item - s_first[wl];
      tl = (int64_t)__ldg(p.word_lists + (size_t)wl * p.n_vol_tiles + (local >> 3));
      sub = (int)(local & 7u);
      tz = (int)(tl % p.tnz); ty = (int)((tl / p.tnz) % p.tny); tx = (int)(tl / ((int64_t)p.tnz * p.tny));
    }
    const int x0 = tx * kTileX, y0 = ty * kTileY, z0 = tz * kTileZ;
    const int lxg = sub >> 2, yh = (sub >> 1) & 1, zh = sub & 1;
    const int lz = (lane & 7) + 8 * zh, ly = (lane >> 3) + 4 * yh;
    // ---- this lane's voxels: kVoxPerThread consecutive x at one (y, z) ------------------------------------------
    const int z = z0 + lz, y = y0 + ly;
    const bool rowok = (z < p.dz) && (y < p.dy);
    const int xb = x0 + lxg * kVoxPerThread;
    unsigned inb = 0u;
    float ptx[kVoxPerThread], pty, ptz;
#pragma unroll
    for (int i = 0; i < kVoxPerThread; ++i) {
      if (rowok && (xb + i < p.dx)) inb |= 1u << i;
      const float vx = (float)(p.xoff + xb + i);
      ptx[i] = SEM == D3M_TSDF_KERNEL_SEMANTICS ? __fmaf_rn(vx, p.vs, p.ox) : __fadd_rn(p.ox, __fmul_rn(p.vs, vx));
    }
    pty = SEM == D3M_TSDF_KERNEL_SEMANTICS ? __fmaf_rn((float)y, p.vs, p.oy) : __fadd_rn(p.oy, __fmul_rn(p.vs, (float)y));
    ptz = SEM == D3M_TSDF_KERNEL_SEMANTICS ? __fmaf_rn(p.vs, (float)z, p.oz) : __fadd_rn(p.oz, __fmul_rn(p.vs, (float)z));
    const int64_t idx0 = ((int64_t)xb * p.dy + y) * p.dz + z;
    const int64_t xstride = (int64_t)p.dy * p.dz;
    float tv[kVoxPerThread], wv[kVoxPerThread], cv[kVoxPerThread];
#pragma unroll
    for (int i = 0; i < kVoxPerThread; ++i) { tv[i] = 1.0f; wv[i] = 0.0f; cv[i] = 0.0f; }
    bool loaded = false;
    unsigned dirty = 0u;
    if (!(p.variant & 4) && !p.direct) {
      // The sub-box's values are requested NOW and first used after the first frame has been projected and its depths
      // sampled: the loads are in flight during that work.  (Loading lazily, at the first update, saved 0.1 GB of DRAM
      // traffic in a kernel that is not DRAM-bound and put a full memory latency in front of the first update: 14 % of
      // the stall samples.)  Tiles are listed only if a frame can reach them, so few of these loads are wasted.
#pragma unroll
      for (int i = 0; i < kVoxPerThread; ++i)
        if ((inb >> i) & 1u) {
          tv[i] = p.tsdf[idx0 + i * xstride]; wv[i] = p.weight[idx0 + i * xstride];
          if (COLOR) cv[i] = p.color[idx0 + i * xstride];
        }
      loaded = true;
    }
    // this warp's sub-box (voxel centres), for the second-level cull
    const int sx0 = x0 + 4 * lxg, sy0 = y0 + 4 * yh, sz0 = z0 + 8 * zh;
    const float wc[3] = {p.ox + ((float)(sx0 + p.xoff) + 1.5f) * p.vs, p.oy + ((float)sy0 + 1.5f) * p.vs,
                         p.oz + ((float)sz0 + 3.5f) * p.vs};
    const float wh[3] = {1.5f * p.vs, 1.5f * p.vs, 3.5f * p.vs};
    // update of the previous frame, applied one pipeline step late (see below)
    unsigned pend_upd = 0u;
    float pend_dist[kVoxPerThread], pend_obs = 0.0f;
    int pend_pix[kVoxPerThread], pend_f = 0;
#pragma unroll
    for (int i = 0; i < kVoxPerThread; ++i) { pend_dist[i] = 0.0f; pend_pix[i] = 0; }
    auto apply_pending = [&]() {
      // (D3M_TSDF_VARIANT & 4) lazy load, per warp: the first frame that updates any voxel of the sub-box brings in its
      // 4 x 8-voxel rows; sub-boxes no frame reaches are never read
      if (!loaded && __any_sync(0xffffffffu, pend_upd != 0u)) {
#pragma unroll
        for (int i = 0; i < kVoxPerThread; ++i)
          if ((inb >> i) & 1u) {
            tv[i] = p.tsdf[idx0 + i * xstride]; wv[i] = p.weight[idx0 + i * xstride];
            if (COLOR) cv[i] = p.color[idx0 + i * xstride];
          }
        loaded = true;
      }
      if (pend_upd == 0u) return;
      const float obs = pend_obs;
      if constexpr (SL) {   // all four quotients, then select
#pragma unroll
        for (int i = 0; i < kVoxPerThread; ++i) {
          const float w_new = __fadd_rn(wv[i], obs);
          const float t_new = __fdiv_rn(__fmaf_rn(pend_dist[i], obs, __fmul_rn(wv[i], tv[i])), w_new);  // :121-126
          const bool on = (pend_upd >> i) & 1u;
          tv[i] = on ? t_new : tv[i];
          wv[i] = on ? w_new : wv[i];
        }
        dirty |= pend_upd;
        pend_upd = 0u;
        return;
      } else {
#pragma unroll
      for (int i = 0; i < kVoxPerThread; ++i) {
        if (!((pend_upd >> i) & 1u)) continue;
        const float w_old = wv[i];
        const float w_new = __fadd_rn(w_old, obs);
        if (SEM == D3M_TSDF_KERNEL_SEMANTICS) {
          tv[i] = __fdiv_rn(__fmaf_rn(pend_dist[i], obs, __fmul_rn(w_old, tv[i])), w_new);  // :121-126
          if (COLOR) {
            // :130-141 (unreachable in the reference because of the `return` at :129); all values are integers < 2^24
            const float oc = cv[i];
            const float ob = floorf(__fmul_rn(oc, 1.0f / 65536.0f));
            const float c0 = __fsub_rn(oc, __fmul_rn(ob, 65536.0f));
            const float og = floorf(__fmul_rn(c0, 1.0f / 256.0f));
            const float orr = __fsub_rn(c0, __fmul_rn(og, 256.0f));
            const float nc = __ldg(p.cimg + (size_t)pend_f * p.H * p.W + pend_pix[i]);
            float nb = floorf(__fmul_rn(nc, 1.0f / 65536.0f));
            const float c1 = __fsub_rn(nc, __fmul_rn(nb, 65536.0f));
            float ng = floorf(__fmul_rn(c1, 1.0f / 256.0f));
            float nr = __fsub_rn(c1, __fmul_rn(ng, 256.0f));
            nb = fminf(roundf(__fdiv_rn(__fmaf_rn(w_old, ob, __fmul_rn(obs, nb)), w_new)), 255.0f);
            ng = fminf(roundf(__fdiv_rn(__fmaf_rn(w_old, og, __fmul_rn(obs, ng)), w_new)), 255.0f);
            nr = fminf(roundf(__fdiv_rn(__fmaf_rn(w_old, orr, __fmul_rn(obs, nr)), w_new)), 255.0f);
            cv[i] = __fadd_rn(__fadd_rn(__fmul_rn(nb, 65536.0f), __fmul_rn(ng, 256.0f)), nr);
          }
        } else {
          tv[i] = __fdiv_rn(__fadd_rn(__fmul_rn(w_old, tv[i]), __fmul_rn(obs, pend_dist[i])), w_new);  // :479
        }
        wv[i] = w_new;
      }
      dirty |= pend_upd;
      pend_upd = 0u;
      }
    };
    // ---- frames in order, one mask word (32 frames) at a time: lane j decides for frame 32w + j whether it can touch
    // THIS sub-box; the warp then walks the surviving frames of the word in order
    for (int w = wl; w < p.words; ++w) {
      const unsigned tmask = p.direct ? (p.F >= 32 ? 0xffffffffu : ((1u << p.F) - 1u)) : __ldg(p.masks + tl * p.words + w);
      if (tmask == 0u) continue;
      const int fmine = 32 * w + lane;
      bool keep = (tmask >> lane) & 1u;
      if (keep && (p.direct || !(p.variant & 1)))
        keep = box_hits_frame(p, fmine, p.direct ? s_zmax32[lane] : __ldg(p.zmax + fmine), wc, wh, use_grid);
      unsigned mask = __ballot_sync(0xffffffffu, keep);
      while (mask) {
        const int j = __ffs(mask) - 1;
        mask &= mask - 1u;
        const int f = 32 * w + j;
        const float* fh = s_hot + f * kHotFloats;
        // software pipeline: project into frame f and issue its depth loads, THEN apply the update of the previous frame
        // (independent arithmetic: two divisions per voxel) while the loads are in flight, then test the depths
        float camz[kVoxPerThread], d[kVoxPerThread];
        int pix[kVoxPerThread];
        const unsigned ok = SL ? project_voxels_sl(p, fh, ptx, pty, ptz, inb, camz, pix)
                               : project_voxels<SEM>(p, fh, ptx, pty, ptz, inb, camz, pix);
        const float* depth = p.depth + (size_t)f * p.H * p.W;
#pragma unroll
        for (int i = 0; i < kVoxPerThread; ++i) d[i] = ((ok >> i) & 1u) ? __ldg(depth + pix[i]) : 0.0f;
        apply_pending();
        pend_upd = SL ? depth_test_voxels_sl(p, ok, d, camz, pend_dist) : depth_test_voxels<SEM>(p, ok, d, camz, pend_dist);
        pend_obs = fh[16];
        if (COLOR) {
          pend_f = f;
#pragma unroll
          for (int i = 0; i < kVoxPerThread; ++i) pend_pix[i] = pix[i];
        }
      }
    }
    apply_pending();
#pragma unroll
    for (int i = 0; i < kVoxPerThread; ++i) {
      if ((dirty >> i) & 1u) {
        p.tsdf[idx0 + i * xstride] = tv[i];
        p.weight[idx0 + i * xstride] = wv[i];
        if (COLOR) p.color[idx0 + i * xstride] = cv[i];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Host-side staging copy.  TSDFVolume.integrate() receives a pageable numpy frame (1.2 MB at 480x640); measured on the
// B200 box a single-threaded memcpy into the pinned ring takes ~130 us and is 90 % of the per-frame cost of the
// reference-shaped API (profiles/r01f_tsdf_host_profile.txt).  A few persistent helper threads split the copy.
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
  int dx, dy, dz, xoff, device, sms, ctas_per_sm;
  float origin[3], vs, trunc;
  float *tsdf, *weight, *color;
  size_t nvox;
  // pinned staging ring for host frames (depth + colour) and their device landing buffers
  float* pinned[kRing];
  float* dframe[kRing];
  cudaEvent_t ev[kRing];
  size_t frame_cap;  // floats per slot (2*H*W)
  int ring_pos;
  // frame tables: per ring slot frames_cap x (kHotFloats + kCullFields) floats, packed for the actual frame count
  float* h_frames[kRing];
  float* d_frames;
  float* d_partial;            // (frames_cap, kFrameAux) coarse depth grids
  float* d_zmax;               // (frames_cap)
  unsigned int* d_counters;    // work queue, word-list lengths
  unsigned int* d_masks;       // (tiles of the volume, frames_cap / 32) frame masks of the last launch
  int* d_word_lists;           // (frames_cap / 32, tiles of the volume)
  int tnx, tny, tnz;
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
  for (int r = 0; r < 3; ++r) {
    for (int c = 0; c < 3; ++c) fr.Wc[4 * r + c] = (float)A[3 * r + c];
    fr.Wc[4 * r + 3] = (float)a[r];
  }
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

static void depth_grid_shape(int H, int W, int& bsl, int& nbx, int& nby) {
  // coarse depth grid: the smallest power-of-two block edge (>= 32 px) that fits kPrepParts x kBlkCols blocks
  bsl = 5;
  while ((((H - 1) >> bsl) + 1 > kPrepParts || ((W - 1) >> bsl) + 1 > kBlkCols) && bsl < 30) ++bsl;
  nbx = ((W - 1) >> bsl) + 1;
  nby = ((H - 1) >> bsl) + 1;
}

// Needs nothing but the depth frames: launched BEFORE the host builds and uploads the frame tables, so that ~60 us of
// host arithmetic per 300 frames (plane equations in fp64) overlap with the 57 us this kernel takes.
static int tsdf_launch_prep(d3m_tsdf* h, const float* depth, int F, int H, int W, cudaStream_t stream) {
  int bsl, nbx, nby;
  depth_grid_shape(H, W, bsl, nbx, nby);
  LaunchScope ls("tsdf_prep", stream);
  tsdf_prep_kernel<<<dim3(kPrepParts, F), 32 * nbx, 0, stream>>>(depth, H, W, bsl, nbx, nby, h->d_partial, h->d_counters,
                                                                kCounters);
  D3M_CUDA_CHECK(cudaGetLastError());
  return D3M_OK;
}

static int tsdf_launch(d3m_tsdf* h, const float* depth, const float* cimg, int F, int H, int W, const float* d_frames,
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
  p.hot = d_frames; p.cull = d_frames + (size_t)F * kHotFloats;
  p.F = F; p.depth = depth; p.cimg = cimg; p.H = H; p.W = W;
  depth_grid_shape(H, W, p.bsl, p.nbx, p.nby);
  p.aux = h->d_partial;
  p.zmax = h->d_zmax;
  p.masks = h->d_masks; p.words = (F + 31) / 32; p.word_lists = h->d_word_lists;
  p.counters = h->d_counters;
  p.tnx = h->tnx; p.tny = h->tny; p.tnz = h->tnz;
  p.n_vol_tiles = (int64_t)h->tnx * h->tny * h->tnz;
  // up to 32 frames (per-frame integrate() calls, the 9 views of a training fragment): one mask word -- skip the cull
  // launch, the integrate kernel tests the frames per sub-box itself (three launches -> two: 35 -> 27 us per call)
  p.direct = (F <= 32 && !(p.variant & 16)) ? 1 : 0;
  if (!p.direct) {
    LaunchScope ls("tsdf_cull", stream);
    tsdf_cull_kernel<<<h->sms * 8, kCullThreads, 0, stream>>>(p);
    D3M_CUDA_CHECK(cudaGetLastError());
  }
  const int sem = flags & 1;
  const bool color = (flags & D3M_TSDF_WITH_COLOR) && cimg != nullptr && sem == D3M_TSDF_KERNEL_SEMANTICS;
  const int grid = h->sms * h->ctas_per_sm;
  LaunchScope ls("tsdf_integrate", stream);
  const size_t smem = sizeof(float) * kHotFloats * (size_t)F;   // <= 40 KB (kMaxFramesPerLaunch)
  if (sem == D3M_TSDF_KERNEL_SEMANTICS) {
    if (color) tsdf_integrate_kernel<D3M_TSDF_KERNEL_SEMANTICS, true, false><<<grid, kTsdfThreads, smem, stream>>>(p);
    else if (p.variant & 8) tsdf_integrate_kernel<D3M_TSDF_KERNEL_SEMANTICS, false, false><<<grid, kTsdfThreads, smem, stream>>>(p);
    else tsdf_integrate_kernel<D3M_TSDF_KERNEL_SEMANTICS, false, true><<<grid, kTsdfThreads, smem, stream>>>(p);
  } else {
    tsdf_integrate_kernel<D3M_TSDF_TORCH_SEMANTICS, false, false><<<grid, kTsdfThreads, smem, stream>>>(p);
  }
  D3M_CUDA_CHECK(cudaGetLastError());
  h->last_launches += p.direct ? 2 : 3;
  return D3M_OK;
}

// Frames per prep -> cull -> integrate round of a batched call.  The depth frames are read twice, by the prep kernel
// (block maxima) and by the integrate kernel; a round whose frames fit in L2 pays DRAM for them once.
// D3M_TSDF_GROUP=<frames> overrides; 0 = as many as one launch can take.
static int frames_per_group(int H, int W) {
  static const int env = getenv("D3M_TSDF_GROUP") ? atoi(getenv("D3M_TSDF_GROUP")) : 0;
  int g = env > 0 ? env : kMaxFramesPerLaunch;
  (void)H; (void)W;
  if (g > kMaxFramesPerLaunch) g = kMaxFramesPerLaunch;
  return g;
}

static int ensure_frames(d3m_tsdf* h, int F) {
  if (F <= h->frames_cap) return D3M_OK;
  int cap = h->frames_cap ? h->frames_cap : 16;
  while (cap < F) cap *= 2;
  if (cap > kMaxFramesPerLaunch) cap = kMaxFramesPerLaunch;
  D3M_CUDA_CHECK(cudaDeviceSynchronize());
  for (int i = 0; i < kRing; ++i) {
    if (h->h_frames[i]) cudaFreeHost(h->h_frames[i]);
    D3M_CUDA_CHECK(cudaMallocHost(&h->h_frames[i], sizeof(float) * (kHotFloats + kCullFields) * (size_t)cap));
  }
  if (h->d_frames) cudaFree(h->d_frames);
  if (h->d_partial) cudaFree(h->d_partial);
  if (h->d_zmax) cudaFree(h->d_zmax);
  if (h->d_masks) cudaFree(h->d_masks);
  if (h->d_word_lists) cudaFree(h->d_word_lists);
  h->d_frames = nullptr; h->d_partial = nullptr; h->d_zmax = nullptr; h->d_masks = nullptr; h->d_word_lists = nullptr;
  D3M_CUDA_CHECK(cudaMalloc(&h->d_frames, sizeof(float) * (kHotFloats + kCullFields) * (size_t)cap * kRing));
  D3M_CUDA_CHECK(cudaMalloc(&h->d_partial, sizeof(float) * (size_t)cap * kFrameAux));
  D3M_CUDA_CHECK(cudaMalloc(&h->d_zmax, sizeof(float) * (size_t)cap));
  D3M_CUDA_CHECK(cudaMalloc(&h->d_masks, sizeof(unsigned int) * (size_t)h->tnx * h->tny * h->tnz * (size_t)((cap + 31) / 32)));
  D3M_CUDA_CHECK(cudaMalloc(&h->d_word_lists, sizeof(int) * (size_t)h->tnx * h->tny * h->tnz * (size_t)((cap + 31) / 32)));
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
  {
    // persistent launch: exactly as many CTAs as fit (the kernel pulls tiles from a queue)
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tsdf_integrate_kernel<D3M_TSDF_KERNEL_SEMANTICS, false, true>,
                                                      kTsdfThreads, 0) != cudaSuccess || occ < 1) occ = 4;
    h->ctas_per_sm = occ;
  }
  cudaError_t e = cudaMalloc(&h->tsdf, h->nvox * 4);
  h->tnx = (dim_x + kTileX - 1) / kTileX; h->tny = (dim_y + kTileY - 1) / kTileY; h->tnz = (dim_z + kTileZ - 1) / kTileZ;
  if (e == cudaSuccess) e = cudaMalloc(&h->d_counters, sizeof(unsigned int) * kCounters);
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
  cudaFree(h->d_frames); cudaFree(h->d_partial); cudaFree(h->d_zmax); cudaFree(h->d_masks); cudaFree(h->d_counters);
  cudaFree(h->d_word_lists);
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

// hot: frame-major (F, kHotFloats); cull: field-major (kCullFields, F)
static void pack_frame(const Frame& fr, float* hot, float* cull, int f, int F) {
  float* q = hot + (size_t)f * kHotFloats;
  q[0] = fr.fx; q[1] = fr.fy; q[2] = fr.cx; q[3] = fr.cy;
  for (int i = 0; i < 12; ++i) q[4 + i] = fr.T[i];
  q[16] = fr.obs; q[17] = 0.f; q[18] = 0.f; q[19] = 0.f;
  auto put = [&](int field, float v) { cull[(size_t)field * F + f] = v; };
  put(kCfFx, fr.fx); put(kCfFy, fr.fy); put(kCfCx, fr.cx); put(kCfCy, fr.cy);
  for (int i = 0; i < 3; ++i) put(kCfCentre + i, fr.centre[i]);
  for (int k = 0; k < 4; ++k) for (int i = 0; i < 3; ++i) put(kCfDirs + 3 * k + i, fr.dirs[k][i]);
  for (int k = 0; k < 5; ++k) for (int i = 0; i < 4; ++i) put(kCfPlanes + 4 * k + i, fr.planes[k][i]);
  for (int i = 0; i < 3; ++i) put(kCfFarN + i, fr.far_n[i]);
  put(kCfFarD0, fr.far_d0);
  for (int i = 0; i < 12; ++i) put(kCfWc + i, fr.Wc[i]);
}

static int upload_frames(d3m_tsdf* h, int F, int H, int W, const float* intr9_host, int intr_per_frame,
                         const float* pose16_host, const float* obs_host, float obs_scalar, int flags,
                         cudaStream_t stream, const float** d_out) {
  int rc = ensure_frames(h, F);
  if (rc != D3M_OK) return rc;
  const int slot = h->fpos;
  h->fpos = (h->fpos + 1) % kRing;
  D3M_CUDA_CHECK(cudaEventSynchronize(h->fev[slot]));  // previous copy out of this pinned table finished
  float* hf = h->h_frames[slot];
  for (int f = 0; f < F; ++f) {
    Frame fr;
    rc = build_frame(fr, intr9_host + (intr_per_frame ? 9 * f : 0), pose16_host + 16 * f,
                     obs_host ? obs_host[f] : obs_scalar, H, W, flags & 1);
    if (rc != D3M_OK) return rc;
    pack_frame(fr, hf, hf + (size_t)F * kHotFloats, f, F);
  }
  const size_t per = (size_t)(kHotFloats + kCullFields);
  float* d = h->d_frames + (size_t)slot * h->frames_cap * per;
  D3M_CUDA_CHECK(cudaMemcpyAsync(d, hf, sizeof(float) * per * F, cudaMemcpyHostToDevice, stream));
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
  const int group = frames_per_group(H, W);
  for (int f0 = 0; f0 < n_frames; f0 += group) {
    const int F = (n_frames - f0) < group ? (n_frames - f0) : group;
    const float* d_frames = nullptr;
    int rc = ensure_frames(h, F);
    if (rc != D3M_OK) return rc;
    rc = tsdf_launch_prep(h, depth + (size_t)f0 * H * W, F, H, W, stream);
    if (rc != D3M_OK) return rc;
    rc = upload_frames(h, F, H, W, intr9_host + (intr_per_frame ? 9 * f0 : 0), intr_per_frame,
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
  const float* d_frames = nullptr;
  int rc = ensure_frames(h, 1);
  if (rc != D3M_OK) return rc;
  rc = tsdf_launch_prep(h, h->dframe[slot], 1, H, W, stream);
  if (rc != D3M_OK) return rc;
  rc = upload_frames(h, 1, H, W, intr9_host, 0, pose16_host, nullptr, obs_weight, flags, stream, &d_frames);
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
