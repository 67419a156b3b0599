// back_project forward for sm_100a: fused voxel->camera projection + in-frustum mask + bilinear gather
// + masked multi-view mean + mean depth, then a deterministic two-kernel depth normalisation.
// Replaces deep3dmap/core/voxel/back_project.py:23-84 of the reference (≈40 aten kernels per fragment).
//
// Work decomposition (one warp = one tile of `tv` consecutive voxels, no block-level sync at all):
//   phase 1  lane <-> (voxel, view slot): a full tile gives every lane one voxel, projected into every view in order;
//            a partial tile (small launches: 16 / 8 / 4 voxels) spreads each voxel's views over 2 / 4 / 8 lanes
//            (push_records).  Either way a 12-byte record {texel offset|corner flags, fx, fy} for every VALID view goes
//            into the warp's shared-memory list (compacted, view order preserved).  Invalid samples cost nothing later.
//   phase 2  lane group (G lanes, each R float4 = 4*G*R channels) <-> voxel: walk the voxel's record list;
//            per record 4*R 128-bit channels-last texel loads per lane (one texel = C contiguous floats), the
//            4-corner FMA chain in the aten order nw,ne,sw,se, then a separate add into the view sum.
//   flush    the tile's (tv, C+1) rows are staged in shared memory and leave with ONE bulk async copy
//            (cp.async.bulk, TMA engine) because (C+1)-float rows cannot be written with aligned vectors.
// Arithmetic order is the oracle's (oracle/d3m_oracle.c) so features and counts are bit-identical to it.
#include <stdlib.h>
#include <string.h>

#include "d3m_common.cuh"

namespace d3m {

constexpr int kFwdWarps = 4;          // warps per CTA (each fully independent)
constexpr int kMaxCountPeers = 16;
constexpr int kMaxViewChunk = 16;     // views whose records are resident in smem at once
constexpr unsigned kFull = 0xffffffffu;
constexpr int kFlagX1 = 1 << 30;      // x0+1 < W
constexpr int kFlagY1 = 1 << 31;      // y0+1 < H
constexpr int kOffMask = (1 << 30) - 1;

struct FwdParams {
  const void* coords;
  int64_t N;
  const float* origin;
  int B;
  float vs;
  const float* feats;
  int V, C, H, W;
  const float* KR;
  float* out;
  float* count;
  float* zbar;
  int* bidx;
  unsigned int* counter;  // ticket of the depth statistics (zero on entry: cleared by the prep kernel, re-cleared by its user)
  int* cell_hist;         // optional histogram of valid samples per (bilinear cell, voxel bucket) bin, consumed by backward
  int nb_log2;            // BinCfg of this call
  // depth statistics fused into this kernel (single-fragment calls): one fp64 (s, s2, n) triple per CTA
  int fuse_stats;
  double* partial;        // (gridDim.x, 3)
  float* stats;           // (B, 2): mean, sd
  double* sums;           // (B, 3) or NULL (voxel-range sharding: handed to the caller instead of finalising)
  int finalize;
  // voxel-range sharding: the view count of every voxel also goes, from this kernel, into every rank's full-scene count
  // buffer at the voxel's global position (peer memory over NVLink) -- the all-gather the next coarse-to-fine level
  // needs, without a collective and without a pass of its own
  int xw, xr;             // world size (0 = off), this rank
  long long xbegin, xblock;   // global row = xbegin + n (contiguous range, xblock == 0) or the block-cyclic map
  float* xcount[kMaxCountPeers];
  int tv, tv_log2;
  unsigned jpat;  // bit k*tv set for every k: the lanes of voxel 0 of a tile (see push_records)
  int vchunk;
  int64_t num_tiles;
  int per_warp_bytes;
};

__device__ __forceinline__ void bulk_store_tile(float* gdst, const float* ssrc, int bytes, int lane) {
  // generic-proxy smem writes -> visible to the async proxy, then one lane hands the tile to the TMA engine
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) {
    const unsigned saddr = (unsigned)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(saddr), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
  __syncwarp();
}

// phase 1 for one chunk of views: the 32 lanes of the warp are (view slot, voxel) pairs -- lane = slot * tv + j holds
// voxel j of the tile and projects views vb + slot of each pass of 32 / tv views, so a 4-voxel tile of the coarse level
// walks its 9 views in 2 dependent projections instead of 9 (a 32-voxel tile: one view per pass, as before).  A valid
// sample's record slot is the voxel's count so far plus its rank among the valid lanes of the same voxel (ballot; the
// lanes of a voxel are in view order), so the records of a voxel stay in ascending view order and the depth sum -- formed
// by the voxel's first lane from the recorded depths, in that order -- keeps the reference's summation order.
// Returns the number of records of voxel j (the same in every lane of the voxel).
template <int KIND>
__device__ __forceinline__ int push_records(const FwdParams& p, int b, int64_t n, float gx, float gy, float gz, int v0,
                                            int v1, int lane, int* rec_off, float* rec_fx, float* rec_fy, float* rec_z,
                                            float& zsum) {
  const float wm1 = (float)(p.W - 1), hm1 = (float)(p.H - 1);
  int ccnt = 0;
  if (p.tv == 32) {
    // full tiles: one lane per voxel, views in sequence (the ballot / recorded-depth bookkeeping below costs the fine
    // level 3 us per launch and buys nothing when every lane already holds a voxel)
    if (b >= 0) {
      for (int v = v0; v < v1; ++v) {
        float4 r0, r1, r2;
        load_krcam(p.KR, v, p.B, b, r0, r1, r2);
        const Sample s = project(gx, gy, gz, r0, r1, r2, wm1, hm1);
        if (s.valid) {
          int off = ((v * p.B + b) * p.H + s.y0) * p.W + s.x0;
          if (p.cell_hist) atomicAdd(p.cell_hist + (((int64_t)off << p.nb_log2) + ((int)n & ((1 << p.nb_log2) - 1))), 1);
          if (s.x0 + 1 < p.W) off |= kFlagX1;
          if (s.y0 + 1 < p.H) off |= kFlagY1;
          rec_off[ccnt * 32 + lane] = off;
          rec_fx[ccnt * 32 + lane] = s.fx;
          rec_fy[ccnt * 32 + lane] = s.fy;
          ++ccnt;
          zsum = __fadd_rn(zsum, s.z);
        }
      }
    }
    return ccnt;
  }
  const int j = lane & (p.tv - 1), slot = lane >> p.tv_log2, vpp = 32 >> p.tv_log2;
  const unsigned jmask = p.jpat << j, below = jmask & ((1u << lane) - 1u);
  for (int vb = v0; vb < v1; vb += vpp) {
    const int v = vb + slot;
    Sample s;
    s.valid = false;
    if (b >= 0 && v < v1) {
      float4 r0, r1, r2;
      load_krcam(p.KR, v, p.B, b, r0, r1, r2);
      s = project(gx, gy, gz, r0, r1, r2, wm1, hm1);
    }
    const unsigned bal = __ballot_sync(kFull, s.valid);
    if (s.valid) {
      int off = ((v * p.B + b) * p.H + s.y0) * p.W + s.x0;
      if (p.cell_hist)  // integer RED: order-independent
        atomicAdd(p.cell_hist + (((int64_t)off << p.nb_log2) + ((int)n & ((1 << p.nb_log2) - 1))), 1);
      if (s.x0 + 1 < p.W) off |= kFlagX1;
      if (s.y0 + 1 < p.H) off |= kFlagY1;
      const int at = (ccnt + __popc(bal & below)) * 32 + j;
      rec_off[at] = off;
      rec_fx[at] = s.fx;
      rec_fy[at] = s.fy;
      rec_z[at] = s.z;
    }
    ccnt += __popc(bal & jmask);
  }
  __syncwarp();
  if (lane < p.tv)
    for (int k = 0; k < ccnt; ++k) zsum = __fadd_rn(zsum, rec_z[k * 32 + lane]);
  return ccnt;
}

__device__ __forceinline__ float corner_chain(float t00, float t01, float t10, float t11, float nw, float ne, float sw,
                                              float se) {
  // aten grid_sampler: out_acc = 0; out_acc += val*w for nw, ne, sw, se  (each contracted to one FMA)
  return __fmaf_rn(t11, se, __fmaf_rn(t10, sw, __fmaf_rn(t01, ne, __fmul_rn(t00, nw))));
}

// mean / sd of back_project.py:77-78 from the three fp64 sums
__device__ __forceinline__ void stats_from_sums(double s, double s2, double c, float& mean, float& sd) {
  if (c > 0.0) {
    mean = (float)(s / c);
    const double m = (double)mean;
    double ssq = s2 - 2.0 * m * s + c * m * m;
    if (ssq < 0.0) ssq = 0.0;
    sd = __fadd_rn((float)sqrt(ssq), 1e-5f);
  } else {
    mean = __int_as_float(0x7fc00000);  // mean of an empty set is NaN in the reference; never used (z<=0 -> 0)
    sd = 1e-5f;
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(kFull, v, o);
  return v;
}

// Depth statistics of a single-fragment call, fused into the gather kernel: every lane sums the mean depths of ITS
// voxels over the warp's tiles in fp64, one shuffle tree per warp, the four warps of the CTA in order, ONE (s, s2, n)
// triple per CTA in global memory.  The triples are folded by whoever needs the result (every CTA of bp_fwd_finish for
// itself: see fold_partials) -- a "last CTA folds" ticket here was measured: the serial fold of a few thousand triples by
// one CTA sits at the tail of the gather kernel (+6 us on every level).  The tile sequence of a warp and both trees are
// pure functions of the launch shape, so the result is bit-identical run to run.
__device__ __forceinline__ void fused_stats_finish(const FwdParams& p, int lane, int warp) {
  __shared__ double s_red[3][kFwdWarps];
  // The sums are formed HERE, from the mean depths this warp stored a moment ago (its own writes, L1/L2-hot), and not
  // inside the tile loop: fp64 accumulators that live across the gather loop cost registers exactly where the loop needs
  // them for loads in flight.
  double s = 0.0, s2 = 0.0, c = 0.0;
  for (int64_t tile = (int64_t)blockIdx.x * kFwdWarps + warp; tile < p.num_tiles;
       tile += (int64_t)gridDim.x * kFwdWarps) {
    const int64_t n = tile * p.tv + lane;
    if (lane < p.tv && n < p.N) {
      const float zb = p.zbar[n];
      if (p.bidx[n] >= 0 && zb > 0.0f) { s += (double)zb; s2 += (double)zb * (double)zb; c += 1.0; }
    }
  }
  s = warp_sum(s); s2 = warp_sum(s2); c = warp_sum(c);
  if (lane == 0) { s_red[0][warp] = s; s_red[1][warp] = s2; s_red[2][warp] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, a2 = 0.0, ac = 0.0;
    for (int w = 0; w < kFwdWarps; ++w) { a += s_red[0][w]; a2 += s_red[1][w]; ac += s_red[2][w]; }
    double* q = p.partial + (size_t)blockIdx.x * 3;
    q[0] = a; q[1] = a2; q[2] = ac;
  }
}

// Fold of the per-CTA triples by one whole CTA of kScanThreads threads: thread t takes triples t, t+256, ... (loads
// batched four deep), one shuffle tree per warp, the warps in order.  A pure function of `nparts`: every CTA that
// calls it gets the same bits.  Result valid in every thread.
__device__ __forceinline__ void fold_partials(const double* __restrict__ partial, int nparts, double& s, double& s2,
                                              double& c) {
  __shared__ double f_red[3][kScanThreads / 32];
  __shared__ double f_out[3];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double a = 0.0, a2 = 0.0, ac = 0.0;
  for (int k0 = tid; k0 < nparts; k0 += 4 * kScanThreads) {
    double v[4][3];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + j * kScanThreads;
      const bool on = k < nparts;
      const double* q = partial + (size_t)(on ? k : 0) * 3;
      v[j][0] = on ? __ldcg(q) : 0.0; v[j][1] = on ? __ldcg(q + 1) : 0.0; v[j][2] = on ? __ldcg(q + 2) : 0.0;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { a += v[j][0]; a2 += v[j][1]; ac += v[j][2]; }
  }
  a = warp_sum(a); a2 = warp_sum(a2); ac = warp_sum(ac);
  if (lane == 0) { f_red[0][warp] = a; f_red[1][warp] = a2; f_red[2][warp] = ac; }
  __syncthreads();
  if (tid == 0) {
    a = 0.0; a2 = 0.0; ac = 0.0;
    for (int w = 0; w < kScanThreads / 32; ++w) { a += f_red[0][w]; a2 += f_red[1][w]; ac += f_red[2][w]; }
    f_out[0] = a; f_out[1] = a2; f_out[2] = ac;
  }
  __syncthreads();
  s = f_out[0]; s2 = f_out[1]; c = f_out[2];
}

template <int KIND, int G, int R>
// 7 CTAs per SM (<= 72 registers) for one float4 per lane, 6 (<= 80) for more: without the bound the partial-tile
// bookkeeping of push_records pushed the level-0 shape to 88 registers (a second wave), with a flat 6 the level-2 shape grew
// from 72 to 80 and lost its seventh CTA (dense forward 187 -> 194 us).
__global__ void __launch_bounds__(kFwdWarps * 32, R == 1 ? 7 : 6) bp_fwd_kernel(const FwdParams p) {

  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NG = 32 / G;  // lane groups per warp
  const int g = lane / G, gl = lane % G;
  const int C = p.C, C4 = C >> 2, C1 = C + 1;
  unsigned char* wbase = smem + (size_t)warp * p.per_warp_bytes;
  float* outs = reinterpret_cast<float*>(wbase);  // (tv, C+1), 16-byte aligned start
  int* rec_off = reinterpret_cast<int*>(wbase + align_up_dev(p.tv * C1 * 4));
  float* rec_fx = reinterpret_cast<float*>(rec_off + p.vchunk * 32);
  float* rec_fy = rec_fx + p.vchunk * 32;
  float* rec_z = rec_fy + p.vchunk * 32;
  const float4* __restrict__ feats4 = reinterpret_cast<const float4*>(p.feats);

  for (int64_t tile = (int64_t)blockIdx.x * kFwdWarps + warp; tile < p.num_tiles;
       tile += (int64_t)gridDim.x * kFwdWarps) {
    const int64_t n0 = tile * p.tv;
    const int64_t n = n0 + (lane & (p.tv - 1));  // lanes tv.. repeat the tile's voxels (view slots of push_records)
    const bool active = lane < p.tv && n < p.N;
    int b = -1;
    float gx = 0.f, gy = 0.f, gz = 0.f;
    if (n < p.N) {
      float cx, cy, cz;
      b = load_coord<KIND>(p.coords, n, p.B, cx, cy, cz);
      if (b >= 0) {
        const float* o = p.origin + 3 * b;
        voxel_world(cx, cy, cz, p.vs, __ldg(o), __ldg(o + 1), __ldg(o + 2), gx, gy, gz);
      }
    }
    int cnt = 0;
    float zsum = 0.0f;
    for (int v0 = 0; v0 < p.V; v0 += p.vchunk) {
      const int v1 = min(p.V, v0 + p.vchunk);
      const bool first = (v0 == 0), last = (v1 == p.V);
      const int ccnt = push_records<KIND>(p, b, n, gx, gy, gz, v0, v1, lane, rec_off, rec_fx, rec_fy, rec_z, zsum);
      cnt += ccnt;
      __syncwarp();
      for (int r = 0; r * NG < p.tv; ++r) {
        const int j = r * NG + g;
        const bool gvalid = (g < NG) && (j < p.tv);
        int cj = __shfl_sync(kFull, ccnt, gvalid ? j : 0);
        const int ctot = __shfl_sync(kFull, cnt, gvalid ? j : 0);
        if (!gvalid) cj = 0;
        const int kmax = __reduce_max_sync(kFull, cj);
        float4 acc[R];
#pragma unroll
        for (int i = 0; i < R; ++i) {
          if (first || !gvalid) {
            acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          } else {
            const float* q = outs + j * C1 + (i * G + gl) * 4;
            acc[i] = make_float4(q[0], q[1], q[2], q[3]);
          }
        }
#pragma unroll 2
        for (int k = 0; k < kmax; ++k) {
          const bool on = k < cj;
          const int jj = on ? j : 0;
          const int of = rec_off[k * 32 + jj];
          const float fx = on ? rec_fx[k * 32 + jj] : 0.0f, fy = on ? rec_fy[k * 32 + jj] : 0.0f;
          const bool x1 = on && (of & kFlagX1), y1 = on && (of & kFlagY1);
          const float wx0 = __fsub_rn(1.0f, fx), wy0 = __fsub_rn(1.0f, fy);
          const float nw = __fmul_rn(wx0, wy0), ne = __fmul_rn(fx, wy0), sw = __fmul_rn(wx0, fy),
                      se = __fmul_rn(fx, fy);
          const float4* base = feats4 + (int64_t)(of & kOffMask) * C4 + gl;
          const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
          float4 t00[R], t01[R], t10[R], t11[R];
#pragma unroll
          for (int i = 0; i < R; ++i) {
            t00[i] = on ? __ldg(base + i * G) : zero;
            t01[i] = x1 ? __ldg(base + C4 + i * G) : zero;
            t10[i] = y1 ? __ldg(base + (int64_t)p.W * C4 + i * G) : zero;
            t11[i] = (x1 && y1) ? __ldg(base + (int64_t)(p.W + 1) * C4 + i * G) : zero;
          }
#pragma unroll
          for (int i = 0; i < R; ++i) {
            acc[i].x = __fadd_rn(acc[i].x, corner_chain(t00[i].x, t01[i].x, t10[i].x, t11[i].x, nw, ne, sw, se));
            acc[i].y = __fadd_rn(acc[i].y, corner_chain(t00[i].y, t01[i].y, t10[i].y, t11[i].y, nw, ne, sw, se));
            acc[i].z = __fadd_rn(acc[i].z, corner_chain(t00[i].z, t01[i].z, t10[i].z, t11[i].z, nw, ne, sw, se));
            acc[i].w = __fadd_rn(acc[i].w, corner_chain(t00[i].w, t01[i].w, t10[i].w, t11[i].w, nw, ne, sw, se));
          }
        }
        if (gvalid) {
          const float div = (float)max(ctot, 1);
#pragma unroll
          for (int i = 0; i < R; ++i) {
            float4 a = acc[i];
            if (last) {  // features /= max(count,1)   (back_project.py:69-72)
              a.x = __fdiv_rn(a.x, div); a.y = __fdiv_rn(a.y, div);
              a.z = __fdiv_rn(a.z, div); a.w = __fdiv_rn(a.w, div);
            }
            float* q = outs + j * C1 + (i * G + gl) * 4;
            q[0] = a.x; q[1] = a.y; q[2] = a.z; q[3] = a.w;
          }
        }
      }
      __syncwarp();
    }
    // per-voxel scalars: count, mean depth (normalised later), fragment index
    if (lane < p.tv) {
      const float zb = __fdiv_rn(zsum, (float)max(cnt, 1));
      outs[lane * C1 + C] = zb;
      if (active) {
        p.count[n] = (float)cnt;
        p.zbar[n] = zb;
        p.bidx[n] = b;
        if (p.xw) {
          const long long gidx = p.xblock > 0 ? ((n / p.xblock) * p.xw + p.xr) * p.xblock + n % p.xblock : p.xbegin + n;
          for (int r = 0; r < p.xw; ++r) p.xcount[r][gidx] = (float)cnt;
        }
      }
    }
    const int rows = (int)min((int64_t)p.tv, p.N - n0);
    const int bytes = rows * C1 * 4;
    float* gdst = p.out + n0 * C1;
    if ((bytes & 15) == 0) {
      bulk_store_tile(gdst, outs, bytes, lane);
    } else {
      __syncwarp();
      for (int i = lane; i < rows * C1; i += 32) gdst[i] = outs[i];
      __syncwarp();
    }
  }
  if (p.fuse_stats) fused_stats_finish(p, lane, warp);
}

// ---- 256-bit variant (C % 8 == 0): a lane owns R groups of EIGHT consecutive channels and fetches each corner of a texel
// with one LDG.E.256 (sm_100: ld.global.v8.f32; a texel is C*4 = 96 / 160 / 320 bytes = 3 / 5 / 10 x 32 bytes and starts on
// a 32-byte boundary).  Half the load instructions of the 128-bit kernel for the same bytes, and twice as many voxels per
// lane-group round (G lanes per voxel instead of 2G).  Same arithmetic per channel, hence the same bits.
struct F8 { float v[8]; };
__device__ __forceinline__ F8 ldg256(const float* p) {
  F8 r;
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
               : "l"(p));
  return r;
}

template <int KIND, int G, int R>
__global__ void __launch_bounds__(kFwdWarps * 32, 6) bp_fwd8_kernel(const FwdParams p) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int NG = 32 / G;  // lane groups per warp
  const int g = lane / G, gl = lane % G;
  const int C = p.C, C1 = C + 1;
  unsigned char* wbase = smem + (size_t)warp * p.per_warp_bytes;
  float* outs = reinterpret_cast<float*>(wbase);  // (tv, C+1), 16-byte aligned start
  int* rec_off = reinterpret_cast<int*>(wbase + align_up_dev(p.tv * C1 * 4));
  float* rec_fx = reinterpret_cast<float*>(rec_off + p.vchunk * 32);
  float* rec_fy = rec_fx + p.vchunk * 32;
  float* rec_z = rec_fy + p.vchunk * 32;

  for (int64_t tile = (int64_t)blockIdx.x * kFwdWarps + warp; tile < p.num_tiles;
       tile += (int64_t)gridDim.x * kFwdWarps) {
    const int64_t n0 = tile * p.tv;
    const int64_t n = n0 + (lane & (p.tv - 1));  // lanes tv.. repeat the tile's voxels (view slots of push_records)
    const bool active = lane < p.tv && n < p.N;
    int b = -1;
    float gx = 0.f, gy = 0.f, gz = 0.f;
    if (n < p.N) {
      float cx, cy, cz;
      b = load_coord<KIND>(p.coords, n, p.B, cx, cy, cz);
      if (b >= 0) {
        const float* o = p.origin + 3 * b;
        voxel_world(cx, cy, cz, p.vs, __ldg(o), __ldg(o + 1), __ldg(o + 2), gx, gy, gz);
      }
    }
    int cnt = 0;
    float zsum = 0.0f;
    for (int v0 = 0; v0 < p.V; v0 += p.vchunk) {
      const int v1 = min(p.V, v0 + p.vchunk);
      const bool first = (v0 == 0), last = (v1 == p.V);
      const int ccnt = push_records<KIND>(p, b, n, gx, gy, gz, v0, v1, lane, rec_off, rec_fx, rec_fy, rec_z, zsum);
      cnt += ccnt;
      __syncwarp();
      for (int r = 0; r * NG < p.tv; ++r) {
        const int j = r * NG + g;
        const bool gvalid = (g < NG) && (j < p.tv);
        int cj = __shfl_sync(kFull, ccnt, gvalid ? j : 0);
        const int ctot = __shfl_sync(kFull, cnt, gvalid ? j : 0);
        if (!gvalid) cj = 0;
        const int kmax = __reduce_max_sync(kFull, cj);
        float acc[R][8];
#pragma unroll
        for (int i = 0; i < R; ++i) {
          const float* q = outs + j * C1 + (i * G + gl) * 8;
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[i][e] = (first || !gvalid) ? 0.0f : q[e];
        }
#pragma unroll 1
        for (int k = 0; k < kmax; ++k) {
          const bool on = k < cj;
          const int jj = on ? j : 0;
          const int of = rec_off[k * 32 + jj];
          const float fx = on ? rec_fx[k * 32 + jj] : 0.0f, fy = on ? rec_fy[k * 32 + jj] : 0.0f;
          const bool x1 = on && (of & kFlagX1), y1 = on && (of & kFlagY1);
          const float wx0 = __fsub_rn(1.0f, fx), wy0 = __fsub_rn(1.0f, fy);
          const float nw = __fmul_rn(wx0, wy0), ne = __fmul_rn(fx, wy0), sw = __fmul_rn(wx0, fy),
                      se = __fmul_rn(fx, fy);
          const float* base = p.feats + (int64_t)(of & kOffMask) * C + gl * 8;
          F8 t00[R], t01[R], t10[R], t11[R];
          F8 zero;
#pragma unroll
          for (int e = 0; e < 8; ++e) zero.v[e] = 0.0f;
#pragma unroll
          for (int i = 0; i < R; ++i) {
            t00[i] = on ? ldg256(base + i * G * 8) : zero;
            t01[i] = x1 ? ldg256(base + C + i * G * 8) : zero;
            t10[i] = y1 ? ldg256(base + (int64_t)p.W * C + i * G * 8) : zero;
            t11[i] = (x1 && y1) ? ldg256(base + (int64_t)(p.W + 1) * C + i * G * 8) : zero;
          }
#pragma unroll
          for (int i = 0; i < R; ++i)
#pragma unroll
            for (int e = 0; e < 8; ++e)
              acc[i][e] = __fadd_rn(acc[i][e], corner_chain(t00[i].v[e], t01[i].v[e], t10[i].v[e], t11[i].v[e], nw, ne, sw, se));
        }
        if (gvalid) {
          const float div = (float)max(ctot, 1);
#pragma unroll
          for (int i = 0; i < R; ++i) {
            float* q = outs + j * C1 + (i * G + gl) * 8;
#pragma unroll
            for (int e = 0; e < 8; ++e) q[e] = last ? __fdiv_rn(acc[i][e], div) : acc[i][e];   // back_project.py:69-72
          }
        }
      }
      __syncwarp();
    }
    // per-voxel scalars: count, mean depth (normalised later), fragment index
    if (lane < p.tv) {
      const float zb = __fdiv_rn(zsum, (float)max(cnt, 1));
      outs[lane * C1 + C] = zb;
      if (active) {
        p.count[n] = (float)cnt;
        p.zbar[n] = zb;
        p.bidx[n] = b;
        if (p.xw) {
          const long long gidx = p.xblock > 0 ? ((n / p.xblock) * p.xw + p.xr) * p.xblock + n % p.xblock : p.xbegin + n;
          for (int r = 0; r < p.xw; ++r) p.xcount[r][gidx] = (float)cnt;
        }
      }
    }
    const int rows = (int)min((int64_t)p.tv, p.N - n0);
    const int bytes = rows * C1 * 4;
    float* gdst = p.out + n0 * C1;
    if ((bytes & 15) == 0) {
      bulk_store_tile(gdst, outs, bytes, lane);
    } else {
      __syncwarp();
      for (int i = lane; i < rows * C1; i += 32) gdst[i] = outs[i];
      __syncwarp();
    }
  }
  if (p.fuse_stats) fused_stats_finish(p, lane, warp);
}

// Any channel count up to 256: the warp takes one voxel at a time, lanes stride over channels (scalar loads).
constexpr int kGenericMaxR = 8;
template <int KIND>
__global__ void __launch_bounds__(kFwdWarps * 32) bp_fwd_generic_kernel(const FwdParams p) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int C = p.C, C1 = C + 1;
  unsigned char* wbase = smem + (size_t)warp * p.per_warp_bytes;
  float* outs = reinterpret_cast<float*>(wbase);
  int* rec_off = reinterpret_cast<int*>(wbase + align_up_dev(p.tv * C1 * 4));
  float* rec_fx = reinterpret_cast<float*>(rec_off + p.vchunk * 32);
  float* rec_fy = rec_fx + p.vchunk * 32;
  float* rec_z = rec_fy + p.vchunk * 32;

  for (int64_t tile = (int64_t)blockIdx.x * kFwdWarps + warp; tile < p.num_tiles;
       tile += (int64_t)gridDim.x * kFwdWarps) {
    const int64_t n0 = tile * p.tv;
    const int64_t n = n0 + (lane & (p.tv - 1));  // lanes tv.. repeat the tile's voxels (view slots of push_records)
    const bool active = lane < p.tv && n < p.N;
    int b = -1;
    float gx = 0.f, gy = 0.f, gz = 0.f;
    if (n < p.N) {
      float cx, cy, cz;
      b = load_coord<KIND>(p.coords, n, p.B, cx, cy, cz);
      if (b >= 0) {
        const float* o = p.origin + 3 * b;
        voxel_world(cx, cy, cz, p.vs, __ldg(o), __ldg(o + 1), __ldg(o + 2), gx, gy, gz);
      }
    }
    int cnt = 0;
    float zsum = 0.0f;
    for (int v0 = 0; v0 < p.V; v0 += p.vchunk) {
      const int v1 = min(p.V, v0 + p.vchunk);
      const bool first = (v0 == 0), last = (v1 == p.V);
      const int ccnt = push_records<KIND>(p, b, n, gx, gy, gz, v0, v1, lane, rec_off, rec_fx, rec_fy, rec_z, zsum);
      cnt += ccnt;
      __syncwarp();
      for (int j = 0; j < p.tv; ++j) {
        const int cj = __shfl_sync(kFull, ccnt, j);
        const int ctot = __shfl_sync(kFull, cnt, j);
        float acc[kGenericMaxR];
#pragma unroll
        for (int i = 0; i < kGenericMaxR; ++i) {
          const int c = lane + 32 * i;
          acc[i] = (first || c >= C) ? 0.0f : outs[j * C1 + c];
        }
        for (int k = 0; k < cj; ++k) {
          const int of = rec_off[k * 32 + j];
          const float fx = rec_fx[k * 32 + j], fy = rec_fy[k * 32 + j];
          const bool x1 = of & kFlagX1, y1 = of & kFlagY1;
          const float wx0 = __fsub_rn(1.0f, fx), wy0 = __fsub_rn(1.0f, fy);
          const float nw = __fmul_rn(wx0, wy0), ne = __fmul_rn(fx, wy0), sw = __fmul_rn(wx0, fy),
                      se = __fmul_rn(fx, fy);
          const float* base = p.feats + (int64_t)(of & kOffMask) * C;
#pragma unroll
          for (int i = 0; i < kGenericMaxR; ++i) {
            const int c = lane + 32 * i;
            if (c < C) {
              const float t00 = __ldg(base + c);
              const float t01 = x1 ? __ldg(base + C + c) : 0.0f;
              const float t10 = y1 ? __ldg(base + (int64_t)p.W * C + c) : 0.0f;
              const float t11 = (x1 && y1) ? __ldg(base + (int64_t)(p.W + 1) * C + c) : 0.0f;
              acc[i] = __fadd_rn(acc[i], corner_chain(t00, t01, t10, t11, nw, ne, sw, se));
            }
          }
        }
        const float div = (float)max(ctot, 1);
#pragma unroll
        for (int i = 0; i < kGenericMaxR; ++i) {
          const int c = lane + 32 * i;
          if (c < C) outs[j * C1 + c] = last ? __fdiv_rn(acc[i], div) : acc[i];
        }
      }
      __syncwarp();
    }
    if (lane < p.tv) {
      const float zb = __fdiv_rn(zsum, (float)max(cnt, 1));
      outs[lane * C1 + C] = zb;
      if (active) {
        p.count[n] = (float)cnt;
        p.zbar[n] = zb;
        p.bidx[n] = b;
        if (p.xw) {
          const long long gidx = p.xblock > 0 ? ((n / p.xblock) * p.xw + p.xr) * p.xblock + n % p.xblock : p.xbegin + n;
          for (int r = 0; r < p.xw; ++r) p.xcount[r][gidx] = (float)cnt;
        }
      }
    }
    const int rows = (int)min((int64_t)p.tv, p.N - n0);
    const int bytes = rows * C1 * 4;
    float* gdst = p.out + n0 * C1;
    if ((bytes & 15) == 0) {
      bulk_store_tile(gdst, outs, bytes, lane);
    } else {
      __syncwarp();
      for (int i = lane; i < rows * C1; i += 32) gdst[i] = outs[i];
      __syncwarp();
    }
  }
  if (p.fuse_stats) fused_stats_finish(p, lane, warp);
}

// ------------------------------------------------------------------------------------------------
// Depth normalisation (back_project.py:77-80): per fragment, over voxels with mean depth > 0,
//   mu = mean(z), sd = ||z - mu||_2 + 1e-5, zn = (z - mu)/sd, zn[z<=0] = 0.
// Deterministic: fixed chunk->CTA assignment, fixed-shape fp64 trees, the last CTA (ticket counter)
// folds the per-chunk partials in chunk order.  Sum(z-mu)^2 is evaluated as Sz2 - 2 mu Sz + n mu^2 in
// fp64 (relative error ~1e-15 here), so one pass over the compact z array suffices.
// ------------------------------------------------------------------------------------------------
constexpr int kStatsThreads = 256;
constexpr int kStatsChunk = 2048;
constexpr int kStatsMaxChunks = 2048;

struct StatsParams {
  const float* zbar;
  const int* bidx;
  int64_t N;
  int B;
  int nchunks;
  int64_t chunk;
  double* partial;  // (nchunks, B, 3)
  float* stats;     // (B, 2): mean, sd
  unsigned int* counter;
  double* sums;     // (B, 3) Sz, Sz2, n+ of this call's voxels, or NULL (voxel-range sharding: all-reduced by the caller)
  int finalize;     // 1: derive stats from this call's sums alone (unsharded)
};

__global__ void __launch_bounds__(kStatsThreads) bp_stats_kernel(const StatsParams p) {
  pdl_enter();
  __shared__ double red[3][kStatsThreads / 32];
  __shared__ int s_bmin, s_bmax;
  __shared__ bool s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t i0 = (int64_t)blockIdx.x * p.chunk, i1 = min(p.N, i0 + p.chunk);
  if (tid == 0) { s_bmin = 0x7fffffff; s_bmax = -1; }
  __syncthreads();
  int bmin = 0, bmax = 0;
  if (p.B > 1) {  // which fragments occur in this chunk (usually one)
    bmin = 0x7fffffff; bmax = -1;
    for (int64_t i = i0 + 4 * tid; i < i1; i += 4 * kStatsThreads) {  // i0 and chunk are multiples of 4
      int bb[4] = {-1, -1, -1, -1};
      if (i + 4 <= i1) {
        const int4 t = *reinterpret_cast<const int4*>(p.bidx + i);
        bb[0] = t.x; bb[1] = t.y; bb[2] = t.z; bb[3] = t.w;
      } else {
        for (int k = 0; k < 4; ++k) if (i + k < i1) bb[k] = p.bidx[i + k];
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) if (bb[k] >= 0) { bmin = min(bmin, bb[k]); bmax = max(bmax, bb[k]); }
    }
    bmin = __reduce_min_sync(kFull, bmin);
    bmax = __reduce_max_sync(kFull, bmax);
    if (lane == 0) { atomicMin(&s_bmin, bmin); atomicMax(&s_bmax, bmax); }
    __syncthreads();
    bmin = s_bmin; bmax = s_bmax;
  }
  double* my = p.partial + (size_t)blockIdx.x * p.B * 3;
  for (int b = tid; b < p.B; b += kStatsThreads)
    if (b < bmin || b > bmax) { my[b * 3 + 0] = 0.0; my[b * 3 + 1] = 0.0; my[b * 3 + 2] = 0.0; }
  for (int b = bmin; b <= bmax; ++b) {
    double s = 0.0, s2 = 0.0, c = 0.0;
    for (int64_t i = i0 + 4 * tid; i < i1; i += 4 * kStatsThreads) {
      float zz[4] = {0.f, 0.f, 0.f, 0.f};
      int bb[4] = {-1, -1, -1, -1};
      if (i + 4 <= i1) {
        const float4 tz = *reinterpret_cast<const float4*>(p.zbar + i);
        const int4 tb = *reinterpret_cast<const int4*>(p.bidx + i);
        zz[0] = tz.x; zz[1] = tz.y; zz[2] = tz.z; zz[3] = tz.w;
        bb[0] = tb.x; bb[1] = tb.y; bb[2] = tb.z; bb[3] = tb.w;
      } else {
        for (int k = 0; k < 4; ++k) if (i + k < i1) { zz[k] = p.zbar[i + k]; bb[k] = p.bidx[i + k]; }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k)  // element order within a thread is fixed -> deterministic
        if (bb[k] == b && zz[k] > 0.0f) { s += (double)zz[k]; s2 += (double)zz[k] * (double)zz[k]; c += 1.0; }
    }
    s = warp_sum(s); s2 = warp_sum(s2); c = warp_sum(c);
    if (lane == 0) { red[0][warp] = s; red[1][warp] = s2; red[2][warp] = c; }
    __syncthreads();
    if (tid == 0) {
      double a = 0.0, a2 = 0.0, ac = 0.0;
      for (int w = 0; w < kStatsThreads / 32; ++w) { a += red[0][w]; a2 += red[1][w]; ac += red[2][w]; }
      my[b * 3 + 0] = a; my[b * 3 + 1] = a2; my[b * 3 + 2] = ac;
    }
    __syncthreads();
  }
  // ticket: the last CTA to arrive folds all partials in chunk order (order-independent of who is last)
  __threadfence();
  if (tid == 0) s_last = (atomicAdd(p.counter, 1u) == (unsigned)(gridDim.x - 1));
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // fixed-shape fold: warp w takes fragments w, w+8, ...; lane l sums chunks l, l+32, ... in order, then one
  // shuffle tree -- the shape depends on nchunks only, never on which CTA happened to arrive last
  for (int b = warp; b < p.B; b += kStatsThreads / 32) {
    double s = 0.0, s2 = 0.0, c = 0.0;
    for (int k = lane; k < p.nchunks; k += 32) {
      const double* q = p.partial + ((size_t)k * p.B + b) * 3;
      s += __ldcg(q); s2 += __ldcg(q + 1); c += __ldcg(q + 2);
    }
    s = warp_sum(s); s2 = warp_sum(s2); c = warp_sum(c);
    if (lane != 0) continue;
    if (p.sums) { p.sums[3 * b] = s; p.sums[3 * b + 1] = s2; p.sums[3 * b + 2] = c; }
    if (p.finalize) {
      float mean, sd;
      stats_from_sums(s, s2, c, mean, sd);
      p.stats[2 * b] = mean;
      p.stats[2 * b + 1] = sd;
    }
  }
  if (tid == 0) *p.counter = 0u;  // ready for the next user of the ticket
}

// voxel-range sharding: stats from the all-reduced sums of every shard
__global__ void bp_stats_from_sums_kernel(const double* __restrict__ sums, float* __restrict__ stats, int B) {
  pdl_enter();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float mean, sd;
  stats_from_sums(sums[3 * b], sums[3 * b + 1], sums[3 * b + 2], mean, sd);
  stats[2 * b] = mean;
  stats[2 * b + 1] = sd;
}

// ------------------------------------------------------------------------------------------------
// bp_fwd_finish: ONE launch for the two independent jobs that follow the gather --
//   CTAs [0, scan_ctas)  exclusive scan of the bin histogram the gather just produced (first step of the backward
//                        pass: it needs nothing but the histogram, so it runs here instead of opening the backward chain);
//   the other CTAs       depth normalisation back_project.py:79-80: out[n, C] = (z - mu) / sd, 0 where z <= 0.
// Either job may be absent (no histogram wanted / sharded call whose statistics are not final yet).
// ------------------------------------------------------------------------------------------------
struct FinishParams {
  BinState bins;
  int scan_ctas;
  int sums_ctas;          // 1: one extra CTA folds the per-CTA triples into `sums` (voxel-range sharding), else 0
  const double* partial;  // per-CTA triples of the gather launch (single-fragment calls), or NULL
  int nparts;             // number of triples; 0 -> `stats` was written by bp_stats_kernel / bp_stats_from_sums_kernel
  double* sums;
  const float* zbar;
  const int* bidx;
  const float* stats;
  float* out;
  int64_t N;
  int C1;
};

__global__ void __launch_bounds__(kScanThreads) bp_fwd_finish_kernel(const FinishParams p) {
  pdl_enter();
  int i = blockIdx.x;
  if (i < p.scan_ctas) {
    scan_bins_cta(p.bins);
    return;
  }
  i -= p.scan_ctas;
  if (i < p.sums_ctas) {
    double s, s2, c;
    fold_partials(p.partial, p.nparts, s, s2, c);
    if (threadIdx.x == 0) { p.sums[0] = s; p.sums[1] = s2; p.sums[2] = c; }
    return;
  }
  i -= p.sums_ctas;
  float mean0 = 0.0f, sd0 = 1.0f;
  if (p.nparts > 0) {  // single fragment: statistics from the gather's per-CTA triples, folded by this CTA for itself
    double s, s2, c;
    fold_partials(p.partial, p.nparts, s, s2, c);
    stats_from_sums(s, s2, c, mean0, sd0);
  }
  const int64_t n = (int64_t)i * kScanThreads + threadIdx.x;
  if (n >= p.N) return;
  const int b = p.bidx[n];
  const float z = p.zbar[n];
  float zn = 0.0f;
  if (b >= 0 && z > 0.0f) {
    const float mean = p.nparts > 0 ? mean0 : p.stats[2 * b], sd = p.nparts > 0 ? sd0 : p.stats[2 * b + 1];
    zn = __fdiv_rn(__fsub_rn(z, mean), sd);
  }
  p.out[n * p.C1 + (p.C1 - 1)] = zn;
}

// ------------------------------------------------------------------------------------------------
// bp_prep: ONE launch in front of the gather --
//   (n_maps, C, H*W) -> (n_maps, H*W, C) relayout of the feature maps through a padded 32x32 shared-memory tile (both
//   sides coalesced; skipped when the producer already is channels-last), and
//   the clear of the binning-state prefix / the statistics ticket (BinLayout), spread over the same CTAs.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bp_prep_kernel(const float* __restrict__ src, float* __restrict__ dst, int A, int Bn,
                                                      uint32_t* __restrict__ zero_ptr, int64_t zero_words) {
  pdl_enter();
  __shared__ float tile[32][33];
  if (zero_words > 0) {
    const int64_t ctas = (int64_t)gridDim.x * gridDim.y * gridDim.z;
    const int64_t cta = ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    const int64_t nv = zero_words >> 2;                       // whole uint4s; the tail words go to CTA 0
    const int64_t per = (nv + ctas - 1) / ctas;
    const int64_t v0 = cta * per, v1 = min(nv, v0 + per);
    uint4* v = reinterpret_cast<uint4*>(zero_ptr);
    for (int64_t i = v0 + threadIdx.x; i < v1; i += 256) v[i] = make_uint4(0u, 0u, 0u, 0u);
    if (cta == 0 && threadIdx.x < (int)(zero_words & 3)) zero_ptr[(nv << 2) + threadIdx.x] = 0u;
  }
  if (src == nullptr) return;
  const int64_t map = blockIdx.z;
  const float* s = src + map * (int64_t)A * Bn;
  float* d = dst + map * (int64_t)A * Bn;
  const int b0 = blockIdx.x * 32, a0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int a = a0 + ty + i, b = b0 + tx;
    if (a < A && b < Bn) tile[ty + i][tx] = __ldg(s + (int64_t)a * Bn + b);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int b = b0 + ty + i, a = a0 + tx;
    if (a < A && b < Bn) d[(int64_t)b * A + a] = tile[tx][ty + i];
  }
}

// Register-transpose variant for C % 4 == 0 and H*W % 4 == 0 (every NeuralRecon level): a thread loads four 128-bit pixel
// quads of four consecutive channels, transposes the 4x4 block in registers and stores four 128-bit channel quads -- no
// shared memory, no barrier, every store instruction of a warp covers whole 32-byte sectors (a pixel's C floats are
// contiguous).  thread <-> (map, pixel quad, channel quad), channel quad fastest.
__global__ void __launch_bounds__(256) bp_prep4_kernel(const float4* __restrict__ src, float4* __restrict__ dst, int C4,
                                                       int HW4, int64_t total, uint32_t* __restrict__ zero_ptr,
                                                       int64_t zero_words) {
  pdl_enter();
  const int64_t nthreads = (int64_t)gridDim.x * 256;
  const int64_t t0 = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (zero_words > 0) {
    const int64_t nv = zero_words >> 2;
    uint4* v = reinterpret_cast<uint4*>(zero_ptr);
    for (int64_t i = t0; i < nv; i += nthreads) v[i] = make_uint4(0u, 0u, 0u, 0u);
    if (t0 < (zero_words & 3)) zero_ptr[(nv << 2) + t0] = 0u;
  }
  for (int64_t t = t0; t < total; t += nthreads) {
    const int cq = (int)(t % C4);
    const int64_t r = t / C4;
    const int pq = (int)(r % HW4);
    const int64_t m = r / HW4;
    const float4* s = src + (m * (4 * C4) + 4 * cq) * HW4 + pq;   // channel 4cq of map m, pixel quad pq
    const float4 a = __ldg(s), b = __ldg(s + HW4), c = __ldg(s + 2 * (int64_t)HW4), d = __ldg(s + 3 * (int64_t)HW4);
    float4* o = dst + (m * (4 * (int64_t)HW4) + 4 * pq) * C4 + cq;  // pixel 4pq of map m, channel quad cq
    o[0] = make_float4(a.x, b.x, c.x, d.x);
    o[C4] = make_float4(a.y, b.y, c.y, d.y);
    o[2 * C4] = make_float4(a.z, b.z, c.z, d.z);
    o[3 * C4] = make_float4(a.w, b.w, c.w, d.w);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct FwdWs {
  size_t zbar, bidx, partial, stats, counter, total;
  int nchunks;
  int64_t chunk;
};

constexpr int kFwdMaxCtasPerSm = 32;
constexpr int kFuseStatsMaxCtas = 1024;
constexpr int kFwdMaxSms = 256;  // bound of the per-warp partial table (a B200 has 148 SMs)

static FwdWs fwd_ws_layout(int64_t N, int B) {
  FwdWs w;
  int64_t nch = (N + kStatsChunk - 1) / kStatsChunk;
  if (nch < 1) nch = 1;
  int64_t chunk = kStatsChunk;
  if (nch > kStatsMaxChunks) {
    chunk = (N + kStatsMaxChunks - 1) / kStatsMaxChunks;
    chunk = (chunk + 3) / 4 * 4;  // vector loads in the stats kernel
    nch = (N + chunk - 1) / chunk;
  }
  w.nchunks = (int)nch;
  w.chunk = chunk;
  // partial table: B > 1 -> (nchunks, B, 3) of the stats kernel; B == 1 -> one triple per CTA of the gather launch
  size_t part = sizeof(double) * 3 * (size_t)nch * (size_t)(B > 0 ? B : 1);
  const size_t fused = sizeof(double) * 3 * (size_t)kFuseStatsMaxCtas;
  if (B == 1 && fused > part) part = fused;
  size_t o = 0;
  w.zbar = o; o = align_up(o + sizeof(float) * (size_t)(N > 0 ? N : 1), 256);
  w.bidx = o; o = align_up(o + sizeof(int) * (size_t)(N > 0 ? N : 1), 256);
  w.partial = o; o = align_up(o + part, 256);
  w.stats = o; o = align_up(o + sizeof(float) * 2 * (size_t)(B > 0 ? B : 1), 256);
  w.counter = o; o = align_up(o + 256, 256);
  w.total = o;
  return w;
}

typedef void (*fwd_kernel_t)(const FwdParams);

static int env_int_early(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && *e) ? atoi(e) : dflt;
}

// (G lanes x R float4 per lane) = C/4 channel quads per voxel.  The table lists every instantiated shape; for a given
// C the FIRST matching row is the default, D3M_FWD_GR="C:G:R[,C:G:R...]" (tuning aid) selects another one.
template <int KIND>
static fwd_kernel_t pick_fwd_kernel(int C, int& G, int& R) {
  struct Row { int g, r; fwd_kernel_t k; };
#define D3M_FWD_ROW(g, r) {g, r, bp_fwd_kernel<KIND, g, r>}
  static const Row rows[] = {
      D3M_FWD_ROW(6, 1),  D3M_FWD_ROW(3, 2),  D3M_FWD_ROW(2, 3),                    // C = 24  (level 2)
      D3M_FWD_ROW(10, 1), D3M_FWD_ROW(5, 2),  D3M_FWD_ROW(2, 5),                    // C = 40  (level 1)
      D3M_FWD_ROW(10, 2), D3M_FWD_ROW(5, 4),  D3M_FWD_ROW(4, 5),                    // C = 80  (level 0)
      D3M_FWD_ROW(4, 1),  D3M_FWD_ROW(8, 1),  D3M_FWD_ROW(16, 1), D3M_FWD_ROW(8, 3), D3M_FWD_ROW(16, 2),
      D3M_FWD_ROW(2, 1),  D3M_FWD_ROW(3, 1),  D3M_FWD_ROW(5, 1)};
#undef D3M_FWD_ROW
  G = 0; R = 0;
  if (C % 4 != 0) return nullptr;
  static const int w8 = env_int_early("D3M_FWD_W8", 0);   // 256-bit texel loads (C % 8 == 0)
  if (w8 && C % 8 == 0) {
    struct Row8 { int g, r; fwd_kernel_t k; };
#define D3M_FWD8_ROW(g, r) {g, r, bp_fwd8_kernel<KIND, g, r>}
    static const Row8 rows8[] = {D3M_FWD8_ROW(3, 1), D3M_FWD8_ROW(5, 1), D3M_FWD8_ROW(10, 1), D3M_FWD8_ROW(5, 2),
                                 D3M_FWD8_ROW(2, 1), D3M_FWD8_ROW(4, 1), D3M_FWD8_ROW(8, 1), D3M_FWD8_ROW(16, 1)};
#undef D3M_FWD8_ROW
    for (const Row8& row : rows8)
      if (row.g * row.r == C / 8 && !(w8 == 2 && C == 80 && row.g == 10)) { G = row.g; R = row.r; return row.k; }
  }
  const int q = C / 4;
  int want_g = 0, want_r = 0;
  if (const char* env = getenv("D3M_FWD_GR")) {
    for (const char* s = env; s && *s;) {
      int c = 0, g = 0, r = 0;
      if (sscanf(s, "%d:%d:%d", &c, &g, &r) == 3 && c == C) { want_g = g; want_r = r; }
      s = strchr(s, ',');
      if (s) ++s;
    }
  }
  const Row* pick = nullptr;
  for (const Row& row : rows) {
    if (row.g * row.r != q) continue;
    if (!pick) pick = &row;
    if (row.g == want_g && row.r == want_r) { pick = &row; break; }
  }
  if (!pick) return nullptr;
  G = pick->g; R = pick->r;
  return pick->k;
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && *e) ? atoi(e) : dflt;
}

template <int KIND>
static int launch_fwd(const FwdParams& p0, cudaStream_t stream, int* grid_out, int* fused_out) {
  FwdParams p = p0;
  int G, R;
  fwd_kernel_t k = pick_fwd_kernel<KIND>(p.C, G, R);
  if (!k) {
    D3M_REQUIRE(p.C <= 32 * kGenericMaxR, D3M_ERR_ARG, "back_project: C=%d unsupported (C%%4!=0 needs C<=%d)", p.C,
                32 * kGenericMaxR);
    k = bp_fwd_generic_kernel<KIND>;
  }
  int sms = current_device_sms();
  if (sms > kFwdMaxSms) sms = kFwdMaxSms;
  // Launch shape.  Large N: 32-voxel warp tiles, grid capped at 32 CTAs per SM (grid-stride).  Small N (the launch is a
  // single wave): the kernel's duration is ONE warp's chain -- views x projection, then tile/NG rounds of count[n] dependent
  // gather steps -- so the tile shrinks, down to D3M_FWD_TVMIN voxels (a multiple of 4 keeps the bulk store 16-byte
  // aligned), until the GPU holds D3M_FWD_WARPS_PER_SM warps per SM.  Measured on the fragment step
  // (profiles/r01j_step_variants.txt): tile floor 4 / 8 / 16 -> level-0 forward 27.7 / 33.5 / 48.2 us; 24 or 32 warps per
  // SM instead of 16 makes level 1 slower (33 vs 28 us).  Tried on top of this and dropped, all bit-identical and all
  // slower or equal: a 4-deep software pipeline of the sample loop, staging the projection rows in shared memory,
  // a flattened (voxel, view) sample list with an in-order shared-memory summation pass (1.5x slower), a 4-deep cp.async
  // ring per lane for the corner texels (1.4x slower) and L1 prefetch of the corners 1-9 samples ahead (2-20 % slower):
  // every extra load/store-unit instruction costs more than the latency it hides.
  static const int tv_min = env_int("D3M_FWD_TVMIN", 4);
  static const int warps_per_sm = env_int("D3M_FWD_WARPS_PER_SM", 16);
  const int tv_floor = tv_min >= 4 && tv_min <= 32 && (tv_min & (tv_min - 1)) == 0 ? tv_min : 8;
  int tv = 32;
  while (tv > tv_floor && (p.N + tv - 1) / tv < (int64_t)sms * warps_per_sm) tv >>= 1;
  p.tv = tv;
  p.tv_log2 = tv == 32 ? 5 : tv == 16 ? 4 : tv == 8 ? 3 : 2;
  p.jpat = tv == 32 ? 1u : 0xffffffffu / ((1u << tv) - 1u);
  p.vchunk = p.V < kMaxViewChunk ? p.V : kMaxViewChunk;
  p.num_tiles = (p.N + tv - 1) / tv;
  p.per_warp_bytes = (int)align_up(align_up((size_t)tv * (p.C + 1) * 4, 16) + (size_t)(tv == 32 ? 3 : 4) * p.vchunk * 32 * 4, 16);   // recorded depths only for partial tiles (push_records)
  const size_t smem = (size_t)p.per_warp_bytes * kFwdWarps;
  D3M_REQUIRE(smem <= 200 * 1024, D3M_ERR_ARG, "back_project: C=%d needs %zu B shared memory per CTA", p.C, smem);
  D3M_CUDA_CHECK(ensure_dynamic_smem(reinterpret_cast<const void*>(k), smem));
  int64_t ctas = (p.num_tiles + kFwdWarps - 1) / kFwdWarps;
  const int64_t cap = (int64_t)sms * kFwdMaxCtasPerSm;
  if (ctas > cap) ctas = cap;
  if (ctas < 1) ctas = 1;
  // Fused statistics only for launches of up to kFuseStatsMaxCtas CTAs: every CTA of bp_fwd_finish folds the per-CTA
  // triples for itself, which is free for a fragment (a few hundred triples from L2) and a 400 MB re-read at dense 96^3
  // (4736 triples x 3456 CTAs: bp_fwd_finish 62 us instead of 21).  Larger launches keep the separate stats kernel.
  if (ctas > kFuseStatsMaxCtas) p.fuse_stats = 0;
  *fused_out = p.fuse_stats;
  *grid_out = (int)ctas;
  {
    LaunchScope ls("bp_fwd", stream);
    launch_k(k, dim3((unsigned)ctas), dim3(kFwdWarps * 32), smem, stream, p);
  }
  D3M_CUDA_CHECK(cudaGetLastError());
  return D3M_OK;
}

// relayout and / or clear in front of the gather (one launch; none when there is nothing to do)
static int launch_prep(const float* src_nchw, float* dst_nhwc, int64_t n_maps, int C, int HW, uint32_t* zero_ptr,
                       int64_t zero_words, cudaStream_t stream) {
  if (src_nchw == nullptr) {
    if (zero_words == 0) return D3M_OK;
    int64_t ctas = ((zero_words >> 2) + 255) / 256;
    if (ctas > 148 * 8) ctas = 148 * 8;
    if (ctas < 1) ctas = 1;
    LaunchScope ls("bp_prep", stream);
    launch_k(bp_prep_kernel, dim3((unsigned)ctas), dim3(256), 0, stream, (const float*)nullptr, (float*)nullptr, 0, 0,
             zero_ptr, zero_words);
    D3M_CUDA_CHECK(cudaGetLastError());
    return D3M_OK;
  }
  static const bool use4 = env_int("D3M_PREP4", 1) != 0;
  if (use4 && (C & 3) == 0 && (HW & 3) == 0 && aligned16(src_nchw) && aligned16(dst_nhwc)) {
    const int64_t total = n_maps * (int64_t)(C / 4) * (HW / 4);
    int64_t ctas = (total + 255) / 256;
    if (ctas > 148 * 32) ctas = 148 * 32;
    LaunchScope ls("bp_prep", stream);
    launch_k(bp_prep4_kernel, dim3((unsigned)ctas), dim3(256), 0, stream, reinterpret_cast<const float4*>(src_nchw),
             reinterpret_cast<float4*>(dst_nhwc), C / 4, HW / 4, total, zero_ptr, zero_words);
    D3M_CUDA_CHECK(cudaGetLastError());
    return D3M_OK;
  }
  for (int64_t m0 = 0; m0 < n_maps; m0 += 65535) {
    const unsigned nz = (unsigned)((n_maps - m0) < 65535 ? (n_maps - m0) : 65535);
    dim3 grid((HW + 31) / 32, (C + 31) / 32, nz);
    LaunchScope ls("bp_prep", stream);
    launch_k(bp_prep_kernel, grid, dim3(256), 0, stream, src_nchw + m0 * (int64_t)C * HW, dst_nhwc + m0 * (int64_t)C * HW, C,
             HW, zero_ptr, m0 == 0 ? zero_words : (int64_t)0);
    D3M_CUDA_CHECK(cudaGetLastError());
  }
  return D3M_OK;
}

}  // namespace d3m

using namespace d3m;

extern "C" size_t d3m_back_project_cell_hist_elems(int64_t N, int B, int V, int H, int W) {
  if (N < 0 || B < 1 || V < 1 || H < 1 || W < 1) return 0;
  return bin_layout(N, B, V, H, W).total;
}

extern "C" size_t d3m_back_project_fwd_workspace(int64_t N, int B, int V, int C) {
  (void)V; (void)C;
  if (N < 0 || B < 1) return 0;
  return fwd_ws_layout(N, B).total;
}

static int fwd_check(const void* coords, int coords_kind, int64_t N, const float* origin, int B, const float* feats,
                     int feats_layout, const float* scratch, int V, int C, int H, int W, const float* KRcam, float* out,
                     float* count, void* workspace, size_t workspace_bytes) {
  D3M_REQUIRE(d3m_device_count() > 0, D3M_ERR_NO_DEVICE, "back_project: no CUDA device (there is no CPU fallback)");
  D3M_REQUIRE(N >= 0 && B >= 1 && V >= 1 && C >= 1 && H >= 2 && W >= 2, D3M_ERR_ARG,
              "back_project: bad sizes N=%lld B=%d V=%d C=%d H=%d W=%d", (long long)N, B, V, C, H, W);
  D3M_REQUIRE(coords_kind >= 0 && coords_kind <= 2, D3M_ERR_ARG, "back_project: coords_kind=%d", coords_kind);
  D3M_REQUIRE(feats_layout == D3M_FEATS_NHWC || feats_layout == D3M_FEATS_NCHW, D3M_ERR_ARG,
              "back_project: feats_layout=%d", feats_layout);
  D3M_REQUIRE((int64_t)V * B * H * W < (1ll << 30), D3M_ERR_ARG, "back_project: V*B*H*W must be < 2^30 texels");
  if (N == 0) return D3M_OK;
  D3M_REQUIRE(coords && origin && feats && KRcam && out && count && workspace, D3M_ERR_ARG,
              "back_project: NULL pointer");
  D3M_REQUIRE(feats_layout == D3M_FEATS_NHWC || scratch, D3M_ERR_ARG,
              "back_project: (V,B,C,H,W) feats need the channels-last scratch buffer");
  D3M_REQUIRE(aligned16(coords) && aligned16(feats) && aligned16(scratch) && aligned16(KRcam) && aligned16(out) &&
                  aligned16(workspace),
              D3M_ERR_ALIGN, "back_project: coords/feats/KRcam/out/workspace must be 16-byte aligned");
  const FwdWs w = fwd_ws_layout(N, B);
  D3M_REQUIRE(workspace_bytes >= w.total, D3M_ERR_WORKSPACE, "back_project: workspace %zu < %zu", workspace_bytes,
              w.total);
  return D3M_OK;
}

// scan of the fresh histogram and / or depth normalisation (and / or the fold of the statistics triples into `sums` for
// the sharded caller), one launch (none when nothing is due)
static int fwd_finish_launch(int64_t N, int C, const FwdWs& w, unsigned char* ws, float* out, int* cell_hist,
                             const BinLayout* bl, bool normalise, int nparts, double* sums, cudaStream_t stream) {
  FinishParams fp;
  memset(&fp, 0, sizeof(fp));
  if (cell_hist) {
    fp.bins = bin_state(cell_hist, *bl);
    fp.scan_ctas = bl->nchunks;
  }
  fp.partial = reinterpret_cast<const double*>(ws + w.partial);
  fp.nparts = nparts;
  fp.sums = sums;
  fp.sums_ctas = (sums && nparts > 0) ? 1 : 0;
  fp.zbar = reinterpret_cast<const float*>(ws + w.zbar);
  fp.bidx = reinterpret_cast<const int*>(ws + w.bidx);
  fp.stats = reinterpret_cast<const float*>(ws + w.stats);
  fp.out = out; fp.N = normalise ? N : 0; fp.C1 = C + 1;
  const int64_t norm_ctas = normalise ? (N + kScanThreads - 1) / kScanThreads : 0;
  const int64_t ctas = fp.scan_ctas + fp.sums_ctas + norm_ctas;
  if (ctas == 0) return D3M_OK;
  LaunchScope ls("bp_fwd_finish", stream);
  launch_k(bp_fwd_finish_kernel, dim3((unsigned)ctas), dim3(kScanThreads), 0, stream, fp);
  D3M_CUDA_CHECK(cudaGetLastError());
  return D3M_OK;
}

// relayout/clear + gather + per-fragment depth sums (+ scan of the histogram, + normalisation);
// `depth_sums` != NULL -> sharded mode: sums are handed to the caller and the depth channel is left un-normalised
// until d3m_back_project_fwd_finish.
static int fwd_impl(const void* coords, int coords_kind, int64_t N, const float* origin, int B, float voxel_size,
                    const float* feats, int feats_layout, float* feats_nhwc_scratch, int V, int C, int H, int W,
                    const float* KRcam, float* out, float* count, int* cell_hist, void* workspace, size_t workspace_bytes,
                    double* depth_sums, cudaStream_t stream, const d3m_count_exchange* cx = nullptr) {
  if (cx) {
    D3M_REQUIRE(cx->peer_count_host && cx->world >= 1 && cx->world <= kMaxCountPeers && cx->rank >= 0 && cx->rank < cx->world &&
                    cx->begin >= 0 && cx->block >= 0,
                D3M_ERR_ARG, "back_project: bad count exchange (at most %d ranks)", kMaxCountPeers);
    for (int r = 0; r < cx->world; ++r)
      D3M_REQUIRE(cx->peer_count_host[r], D3M_ERR_ARG, "back_project: count buffer of rank %d is NULL", r);
  }
  int rc = fwd_check(coords, coords_kind, N, origin, B, feats, feats_layout, feats_nhwc_scratch, V, C, H, W, KRcam, out,
                     count, workspace, workspace_bytes);
  if (rc != D3M_OK) return rc;
  PdlScope pdl_scope(stream, (long long)N * V);
  BinLayout bl;
  memset(&bl, 0, sizeof(bl));
  if (cell_hist) {
    D3M_REQUIRE(aligned16(cell_hist), D3M_ERR_ALIGN, "back_project: cell_hist must be 16-byte aligned");
    bl = bin_layout(N, B, V, H, W);
  }
  if (N == 0) {
    if (cell_hist) D3M_CUDA_CHECK(cudaMemsetAsync(cell_hist, 0, sizeof(int) * bl.total, stream));  // empty bins, start = 0
    if (depth_sums) D3M_CUDA_CHECK(cudaMemsetAsync(depth_sums, 0, sizeof(double) * 3 * (size_t)B, stream));
    return D3M_OK;
  }
  const FwdWs w = fwd_ws_layout(N, B);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  // the statistics ticket lives in the binning state when there is one (cleared together with it), else in the workspace
  unsigned int* ticket = cell_hist ? bin_state(cell_hist, bl).counters + kCtrStatsTicket
                                   : reinterpret_cast<unsigned int*>(ws + w.counter);
  uint32_t* zero_ptr = cell_hist ? reinterpret_cast<uint32_t*>(cell_hist) : reinterpret_cast<uint32_t*>(ws + w.counter);
  // without a binning state only the stats kernel's ticket needs clearing; a single small fragment (fused statistics)
  // does not use it, and is the one case where the clear would be a launch of its own
  const bool maybe_fused = B == 1 && (N + 3) / 4 / kFwdWarps <= kFuseStatsMaxCtas;
  const int64_t zero_words = cell_hist ? (int64_t)bl.zero_elems : (maybe_fused ? 0 : 4);
  const float* feats_nhwc = feats;
  if (feats_layout == D3M_FEATS_NCHW) {
    rc = launch_prep(feats, feats_nhwc_scratch, (int64_t)V * B, C, H * W, zero_ptr, zero_words, stream);
    feats_nhwc = feats_nhwc_scratch;
  } else {
    rc = launch_prep(nullptr, nullptr, 0, 0, 0, zero_ptr, zero_words, stream);
  }
  if (rc != D3M_OK) return rc;
  FwdParams p;
  memset(&p, 0, sizeof(p));
  p.coords = coords; p.N = N; p.origin = origin; p.B = B; p.vs = voxel_size;
  p.feats = feats_nhwc; p.V = V; p.C = C; p.H = H; p.W = W; p.KR = KRcam;
  p.out = out; p.count = count;
  p.cell_hist = cell_hist ? cell_hist + bl.cnt : nullptr;
  p.nb_log2 = cell_hist ? bl.nb_log2 : 0;
  p.zbar = reinterpret_cast<float*>(ws + w.zbar);
  p.bidx = reinterpret_cast<int*>(ws + w.bidx);
  p.counter = ticket;
  p.fuse_stats = (B == 1) ? 1 : 0;
  p.partial = reinterpret_cast<double*>(ws + w.partial);
  p.stats = reinterpret_cast<float*>(ws + w.stats);
  p.sums = depth_sums;
  p.finalize = depth_sums ? 0 : 1;
  if (cx) {
    p.xw = cx->world; p.xr = cx->rank; p.xbegin = cx->begin; p.xblock = cx->block;
    for (int r = 0; r < cx->world; ++r) p.xcount[r] = static_cast<float*>(cx->peer_count_host[r]);
  }
  int fwd_grid = 0, fused = 0;
  if (coords_kind == D3M_COORDS_F32) rc = launch_fwd<D3M_COORDS_F32>(p, stream, &fwd_grid, &fused);
  else if (coords_kind == D3M_COORDS_I64) rc = launch_fwd<D3M_COORDS_I64>(p, stream, &fwd_grid, &fused);
  else rc = launch_fwd<D3M_COORDS_I32>(p, stream, &fwd_grid, &fused);
  p.fuse_stats = fused;
  if (rc != D3M_OK) return rc;
  if (!p.fuse_stats) {
    StatsParams sp;
    sp.zbar = p.zbar; sp.bidx = p.bidx; sp.N = N; sp.B = B; sp.nchunks = w.nchunks; sp.chunk = w.chunk;
    sp.partial = p.partial;
    sp.stats = p.stats;
    sp.counter = ticket;
    sp.sums = depth_sums;
    sp.finalize = depth_sums ? 0 : 1;
    LaunchScope ls("bp_fwd_stats", stream);
    launch_k(bp_stats_kernel, dim3(w.nchunks), dim3(kStatsThreads), 0, stream, sp);
    D3M_CUDA_CHECK(cudaGetLastError());
  }
  return fwd_finish_launch(N, C, w, ws, out, cell_hist, &bl, depth_sums == nullptr, p.fuse_stats ? fwd_grid : 0,
                           p.fuse_stats ? depth_sums : nullptr, stream);
}

extern "C" int d3m_back_project_fwd(const void* coords, int coords_kind, int64_t N, const float* origin, int B,
                                    float voxel_size, const float* feats, int feats_layout, float* feats_nhwc_scratch,
                                    int V, int C, int H, int W, const float* KRcam, float* out, float* count,
                                    int* cell_hist, void* workspace, size_t workspace_bytes, void* stream_) {
  return fwd_impl(coords, coords_kind, N, origin, B, voxel_size, feats, feats_layout, feats_nhwc_scratch, V, C, H, W,
                  KRcam, out, count, cell_hist, workspace, workspace_bytes, nullptr, static_cast<cudaStream_t>(stream_));
}

extern "C" int d3m_back_project_fwd_partial(const void* coords, int coords_kind, int64_t N, const float* origin, int B,
                                            float voxel_size, const float* feats, int feats_layout,
                                            float* feats_nhwc_scratch, int V, int C, int H, int W, const float* KRcam,
                                            float* out, float* count, int* cell_hist, double* depth_sums,
                                            void* workspace, size_t workspace_bytes, void* stream_) {
  D3M_REQUIRE(depth_sums, D3M_ERR_ARG, "back_project_fwd_partial: NULL depth_sums");
  return fwd_impl(coords, coords_kind, N, origin, B, voxel_size, feats, feats_layout, feats_nhwc_scratch, V, C, H, W,
                  KRcam, out, count, cell_hist, workspace, workspace_bytes, depth_sums, static_cast<cudaStream_t>(stream_));
}

extern "C" int d3m_back_project_fwd_finish(int64_t N, int B, int C, const double* depth_sums, float* out,
                                           void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  D3M_REQUIRE(d3m_device_count() > 0, D3M_ERR_NO_DEVICE, "back_project: no CUDA device (there is no CPU fallback)");
  D3M_REQUIRE(N >= 0 && B >= 1 && C >= 1, D3M_ERR_ARG, "back_project_fwd_finish: bad sizes");
  if (N == 0) return D3M_OK;
  D3M_REQUIRE(depth_sums && out && workspace, D3M_ERR_ARG, "back_project_fwd_finish: NULL pointer");
  const FwdWs w = fwd_ws_layout(N, B);
  D3M_REQUIRE(workspace_bytes >= w.total, D3M_ERR_WORKSPACE, "back_project_fwd_finish: workspace %zu < %zu",
              workspace_bytes, w.total);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  {
    LaunchScope ls("bp_fwd_stats_from_sums", stream);
    launch_k(bp_stats_from_sums_kernel, dim3((B + 127) / 128), dim3(128), 0, stream, depth_sums, reinterpret_cast<float*>(ws + w.stats), B);
  }
  D3M_CUDA_CHECK(cudaGetLastError());
  return fwd_finish_launch(N, C, w, ws, out, nullptr, nullptr, true, 0, nullptr, stream);
}

extern "C" int d3m_back_project_fwd_partial_x(const void* coords, int coords_kind, int64_t N, const float* origin, int B,
                                              float voxel_size, const float* feats, int feats_layout,
                                              float* feats_nhwc_scratch, int V, int C, int H, int W, const float* KRcam,
                                              float* out, float* count, int* cell_hist, double* depth_sums,
                                              void* workspace, size_t workspace_bytes, const d3m_count_exchange* cx,
                                              void* stream_) {
  D3M_REQUIRE(depth_sums, D3M_ERR_ARG, "back_project_fwd_partial_x: NULL depth_sums");
  return fwd_impl(coords, coords_kind, N, origin, B, voxel_size, feats, feats_layout, feats_nhwc_scratch, V, C, H, W,
                  KRcam, out, count, cell_hist, workspace, workspace_bytes, depth_sums, static_cast<cudaStream_t>(stream_), cx);
}

