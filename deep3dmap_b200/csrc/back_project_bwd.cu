// back_project backward (grad w.r.t. the 2D feature maps) for sm_100a -- deterministic, no float atomics.
// Replaces autograd through deep3dmap/core/voxel/back_project.py:55-73 of the reference:
//   grad_sample[v,n,:] = grad_out[n,:C] / max(count[n],1)   masked to valid views,
//   grid_sampler_2d_backward: grad_feats[v,b,y,x,:] += w_corner(v,n) * grad_sample[v,n,:]  (aten uses atomicAdd).
//
// The scatter is turned into a gather (texel-centric), so every texel is produced by exactly one lane group
// in a fixed order:
//   hist    histogram of the valid samples per bin = (bilinear cell (v,b,y0,x0), voxel bucket), INTEGER atomics --
//           normally produced by the forward gather, which projects every voxel anyway (d3m_back_project_fwd(cell_hist));
//   scan    exclusive prefix sum of the histogram (single pass, decoupled look-back) -- normally done by the
//           scan CTAs of bp_fwd_finish right after the forward gather;
//   fill    one launch, two independent jobs: project again, claim a slot of the bin with an integer atomic and store
//           {n, fx, fy, bin} (16 B); the spare CTAs write the pre-divided rows ghat[n,:] = grad_out[n,:C]/max(count,1)
//           into a 16-byte aligned (N,C) buffer (grad_out rows are (C+1) floats and cannot be vector-loaded);
//   order   every bin's entries into ascending voxel order (thread per entry, rank sort, out of place) -- this removes
//           the only nondeterminism, the claim order of the fill atomics;
//   gather  CTA per (TY x TX) tile of texels of one map: lane groups walk the (TY+1) x (TX+1) bilinear CELLS that
//           touch the tile; every entry's ghat row is loaded ONCE (R 128-bit loads per lane) and feeds the four
//           corner sums nw/ne/sw/se of its cell (fp32 mul + add in entry order).  The 4 per-cell partials go to
//           shared memory; after one barrier every texel adds its four partials in the fixed order
//           nw(y,x) + ne(y,x-1) + sw(y-1,x) + se(y-1,x-1) and is stored once.  (A texel-centric walk reads every
//           row four times and was bound by L2->SM load latency.)
#include <stdlib.h>
#include <string.h>

#include "d3m_common.cuh"

namespace d3m {

constexpr unsigned kFullB = 0xffffffffu;
constexpr int kGatherWarps = 8;
constexpr int kMaxExchangeRanks = 16;

struct BwdParams {
  const void* coords;
  int64_t N;
  const float* origin;
  int B;
  float vs;
  int V, C, H, W;
  const float* KR;
  const float* grad_out;
  const float* count;  // (N,) view counts from the forward pass, or the workspace copy computed by bp_bwd_count_kernel
  float* ghat;
  int* bin_cnt;          // histogram (BinLayout.cnt): forward's, or this call's workspace copy
  int* bin_cursor;       // claims per bin of the fill pass; zero on entry, re-zeroed by the gather (BinLayout.cursor)
  int* bin_start;        // Mb+1, exclusive scan of bin_cnt
  int4* entries;         // {voxel, fx, fy, bin} in slot-claim order (fill)
  int4* sorted;          // the same entries, every bin in ascending voxel order (gather phase 0)
  float* grad_feats;
  int64_t M;      // V*B*H*W cells
  int64_t Mb;     // M << nb_log2 bins (cell-major, voxel-bucket-minor; see BinCfg)
  int nb_log2;
  int grad_nchw;                   // 1: gather writes (V,B,C,H,W) directly
  // view-owner exchange (voxel-range sharding over several GPUs): the gather stores every texel straight into the staging
  // buffer of the rank that owns the texel's view -- peer memory over NVLink -- in slot `xchg_rank` of that buffer
  int xchg_world;                  // 0 = off
  int xchg_rank, xchg_vpo;         // this rank; views per owner
  float* xchg_peer[kMaxExchangeRanks];  // staging buffer of every rank, (world, vpo, B, H, W, C), mapped in this process
  // fill + ghat launch
  int fill_ctas_x, ghat_ctas;      // voxel CTAs per view group of the fill pass; CTAs of the pre-division pass (they come first)
};

constexpr int kSampleThreads = 256;
constexpr int kGhatSteps = 4;       // row groups in flight per warp of the pre-division pass

// Only used when the caller does not hand over the forward pass's `count`: valid views per voxel.
template <int KIND>
__global__ void __launch_bounds__(kSampleThreads) bp_bwd_count_kernel(const BwdParams p, float* __restrict__ cnt_out) {
  pdl_enter();
  const int64_t n = (int64_t)blockIdx.x * kSampleThreads + threadIdx.x;
  if (n >= p.N) return;
  float cx, cy, cz;
  const int b = load_coord<KIND>(p.coords, n, p.B, cx, cy, cz);
  int cnt = 0;
  if (b >= 0) {
    float gx, gy, gz;
    const float* o = p.origin + 3 * b;
    voxel_world(cx, cy, cz, p.vs, __ldg(o), __ldg(o + 1), __ldg(o + 2), gx, gy, gz);
    const float wm1 = (float)(p.W - 1), hm1 = (float)(p.H - 1);
    for (int v = 0; v < p.V; ++v) {
      float4 r0, r1, r2;
      load_krcam(p.KR, v, p.B, b, r0, r1, r2);
      cnt += project(gx, gy, gz, r0, r1, r2, wm1, hm1).valid ? 1 : 0;
    }
  }
  cnt_out[n] = (float)cnt;
}

// One thread per (voxel, group of kViewsPerThread views).  Short dependency chains and N*V-way parallelism instead of
// a 9-deep serial loop per voxel (these passes are latency-bound at fragment size).
//   FILL = false: histogram the valid samples per bin (integer RED).  Only when the forward pass did not already
//                 produce the histogram (d3m_back_project_fwd(..., cell_hist)).
//   FILL = true : claim the bin's next entry position (one integer atomic on the cursor), store {n, fx, fy, bin}.
constexpr int kViewsPerThread = 3;  // independent atomics in flight per thread (the pass is latency-bound on them)
template <int KIND, bool FILL>
__device__ __forceinline__ void sample_pass(const BwdParams& p, const int64_t n, const int v0) {
  if (n >= p.N) return;
  float cx, cy, cz;
  const int b = load_coord<KIND>(p.coords, n, p.B, cx, cy, cz);
  if (b < 0) return;
  float gx, gy, gz;
  const float* o = p.origin + 3 * b;
  voxel_world(cx, cy, cz, p.vs, __ldg(o), __ldg(o + 1), __ldg(o + 2), gx, gy, gz);
  int key[kViewsPerThread];
  float fx[kViewsPerThread], fy[kViewsPerThread];
  bool ok[kViewsPerThread];
#pragma unroll
  for (int j = 0; j < kViewsPerThread; ++j) {
    const int v = v0 + j;
    ok[j] = v < p.V;
    key[j] = 0; fx[j] = 0.f; fy[j] = 0.f;
    if (ok[j]) {
      float4 r0, r1, r2;
      load_krcam(p.KR, v, p.B, b, r0, r1, r2);
      const Sample s = project(gx, gy, gz, r0, r1, r2, (float)(p.W - 1), (float)(p.H - 1));
      ok[j] = s.valid;
      key[j] = ((((v * p.B + b) * p.H + s.y0) * p.W + s.x0) << p.nb_log2) + ((int)n & ((1 << p.nb_log2) - 1));
      fx[j] = s.fx; fy[j] = s.fy;
    }
  }
  if (!FILL) {
#pragma unroll
    for (int j = 0; j < kViewsPerThread; ++j)
      if (ok[j]) atomicAdd(p.bin_cnt + key[j], 1);
  } else {
    int pos[kViewsPerThread];
#pragma unroll
    for (int j = 0; j < kViewsPerThread; ++j)   // claim order is arbitrary; the gather's phase 0 fixes it
      pos[j] = ok[j] ? __ldg(p.bin_start + key[j]) + atomicAdd(p.bin_cursor + key[j], 1) : 0;
#pragma unroll
    for (int j = 0; j < kViewsPerThread; ++j)
      if (ok[j]) p.entries[pos[j]] = make_int4((int)n, __float_as_int(fx[j]), __float_as_int(fy[j]), key[j]);
  }
}

template <int KIND>
__global__ void __launch_bounds__(kSampleThreads) bp_bwd_hist_kernel(const BwdParams p) {
  pdl_enter();
  sample_pass<KIND, false>(p, (int64_t)blockIdx.x * kSampleThreads + threadIdx.x, blockIdx.y * kViewsPerThread);
}

// ghat[n,c] = grad_out[n,c] / max(count[n],1)   (div backward of back_project.py:72), re-packed to 16-byte aligned
// rows of C floats.  `wg` = global warp index of this job.
__device__ __forceinline__ void ghat_rows(const BwdParams& p, const int64_t wg, const int lane) {
  const int C = p.C, C1 = C + 1;
  if ((C & 3) == 0 && C <= 128) {
    // lane <-> (row, channel quad): 32/(C/4) rows per warp step, kGhatSteps steps in flight; 4 scalar loads (rows of
    // C+1 floats are unaligned), one 128-bit store per lane
    const int C4 = C >> 2;
    const int rows_per_step = 32 / C4;
    const int rl = lane / C4, j = lane - rl * C4;
    const bool lane_on = rl < rows_per_step;
    const int64_t r0 = wg * (int64_t)(rows_per_step * kGhatSteps) + rl;
    float4 g[kGhatSteps];
    float d[kGhatSteps];
#pragma unroll
    for (int k = 0; k < kGhatSteps; ++k) {
      const int64_t r = r0 + (int64_t)k * rows_per_step;
      g[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      d[k] = 1.0f;
      if (lane_on && r < p.N) {
        const float* src = p.grad_out + r * C1 + 4 * j;
        g[k] = make_float4(__ldg(src), __ldg(src + 1), __ldg(src + 2), __ldg(src + 3));
        d[k] = fmaxf(__ldg(p.count + r), 1.0f);
      }
    }
#pragma unroll
    for (int k = 0; k < kGhatSteps; ++k) {
      const int64_t r = r0 + (int64_t)k * rows_per_step;
      if (lane_on && r < p.N)
        reinterpret_cast<float4*>(p.ghat)[r * C4 + j] =
            make_float4(__fdiv_rn(g[k].x, d[k]), __fdiv_rn(g[k].y, d[k]), __fdiv_rn(g[k].z, d[k]), __fdiv_rn(g[k].w, d[k]));
    }
  } else {
    // any C: warp per row, lanes stride over the channels
    const int64_t r0 = wg * kGhatSteps;
    for (int k = 0; k < kGhatSteps; ++k) {
      const int64_t r = r0 + k;
      if (r >= p.N) break;
      const float d = fmaxf(__ldg(p.count + r), 1.0f);
      for (int c = lane; c < C; c += 32) p.ghat[r * C + c] = __fdiv_rn(__ldg(p.grad_out + r * C1 + c), d);
    }
  }
}

// ONE launch, two independent jobs: CTAs [0, ghat_ctas) run the pre-division pass, the CTAs behind them the fill pass
// (CTA j <-> voxel block j % fill_ctas_x, view group j / fill_ctas_x).
template <int KIND>
__global__ void __launch_bounds__(kSampleThreads) bp_bwd_fill_ghat_kernel(const BwdParams p) {
  pdl_enter();
  // the streaming job first: its CTAs are dispatched together with the first fill CTAs, so that its DRAM traffic overlaps
  // the fill pass's atomics instead of trailing it
  const int i = blockIdx.x;
  if (i < p.ghat_ctas) {
    ghat_rows(p, (int64_t)i * (kSampleThreads / 32) + (threadIdx.x >> 5), threadIdx.x & 31);
    return;
  }
  const int j = i - p.ghat_ctas;
  const int bx = j % p.fill_ctas_x, by = j / p.fill_ctas_x;
  sample_pass<KIND, true>(p, (int64_t)bx * kSampleThreads + threadIdx.x, by * kViewsPerThread);
}

// legacy path (no forward histogram): scan as a kernel of its own
struct ScanOnly { BinState bins; };
__global__ void __launch_bounds__(kScanThreads) bp_bwd_scan_kernel(const ScanOnly p) {
  pdl_enter();
  scan_bins_cta(p.bins);
}

// ---- gather --------------------------------------------------------------------------------------
// One batch of up to BS entries of a cell, held one per lane (lane j of the group = entry j, .x < 0 = absent): the
// group issues all BS row loads back to back (BS*R independent 128-bit loads in flight per lane) and then does the
// arithmetic in entry order.  An absent entry has a zero row: it adds +0 to every sum (no branch needed).
// Fused multiply-add: one rounding per contribution (aten rounds the product first; the difference is <= 0.5 ulp
// per term, inside the 1e-5 gradient tolerance) and half the FP instructions of mul + add.
template <int G, int R, int BS>
__device__ __forceinline__ void gather_batch(const int4 cur, const int gbase, const int gl, const bool gact,
                                             const float4* __restrict__ ghat4, float4 (&anw)[R], float4 (&ane)[R],
                                             float4 (&asw)[R], float4 (&ase)[R]) {
  constexpr int C4 = G * R;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 rows[BS][R];
  float fxs[BS], fys[BS];
#pragma unroll
  for (int j = 0; j < BS; ++j) {
    const int nj = __shfl_sync(kFullB, cur.x, gbase + j);
    fxs[j] = __int_as_float(__shfl_sync(kFullB, cur.y, gbase + j));
    fys[j] = __int_as_float(__shfl_sync(kFullB, cur.z, gbase + j));
    const bool on = gact && nj >= 0;
    const float4* row = ghat4 + (int64_t)(on ? nj : 0) * C4 + gl;
#pragma unroll
    for (int i = 0; i < R; ++i) rows[j][i] = on ? __ldg(row + i * G) : zero4;
  }
#pragma unroll
  for (int j = 0; j < BS; ++j) {
    const float fx = fxs[j], fy = fys[j];
    const float wx0 = __fsub_rn(1.0f, fx), wy0 = __fsub_rn(1.0f, fy);
    const float nw = __fmul_rn(wx0, wy0), ne = __fmul_rn(fx, wy0), sw = __fmul_rn(wx0, fy), se = __fmul_rn(fx, fy);
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const float4 q = rows[j][i];
      anw[i].x = __fmaf_rn(nw, q.x, anw[i].x); anw[i].y = __fmaf_rn(nw, q.y, anw[i].y);
      anw[i].z = __fmaf_rn(nw, q.z, anw[i].z); anw[i].w = __fmaf_rn(nw, q.w, anw[i].w);
      ane[i].x = __fmaf_rn(ne, q.x, ane[i].x); ane[i].y = __fmaf_rn(ne, q.y, ane[i].y);
      ane[i].z = __fmaf_rn(ne, q.z, ane[i].z); ane[i].w = __fmaf_rn(ne, q.w, ane[i].w);
      asw[i].x = __fmaf_rn(sw, q.x, asw[i].x); asw[i].y = __fmaf_rn(sw, q.y, asw[i].y);
      asw[i].z = __fmaf_rn(sw, q.z, asw[i].z); asw[i].w = __fmaf_rn(sw, q.w, asw[i].w);
      ase[i].x = __fmaf_rn(se, q.x, ase[i].x); ase[i].y = __fmaf_rn(se, q.y, ase[i].y);
      ase[i].z = __fmaf_rn(se, q.z, ase[i].z); ase[i].w = __fmaf_rn(se, q.w, ase[i].w);
    }
  }
}

__device__ __forceinline__ float4 add4(const float4 a, const float4 b) {
  return make_float4(__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w));
}

// Measured and dropped (profiles/r02x_step_variants.txt): a prologue that fetches every cell's [start, end) range and, if
// they fit, the window's entries into shared memory (two parallel steps instead of a dependent chain per cell and round):
// the extra barriers and the serial row bookkeeping cost more than the chain -- fragment level 2 50 -> 60-65 us, dense
// level 2 113 -> 144 us.
constexpr int kBigCell = 256;  // entries from which a cell is processed by the whole CTA instead of one lane group
constexpr int kMaxBig = 64;    // deferred cells per tile (further ones are simply processed by their own group)

// Shared memory: part[4][ncell][C/4] float4 -- corner-major so that the lane groups of a warp (adjacent cells)
// touch adjacent addresses in both phases (no bank conflicts, no padding) -- then scratch[warps][4][C/4] float4 for
// the cooperative pass over heavily populated cells.
template <int G, int R>
__global__ void __launch_bounds__(kGatherWarps * 32, (R == 1 ? 4 : 3)) bp_bwd_gather_tile_kernel(const BwdParams p, const int TX,
                                                                               const int TY, const int tiles_x,
                                                                               const int tiles_y) {
  pdl_enter();
  extern __shared__ __align__(16) unsigned char gsm[];
  __shared__ int s_big[kMaxBig];
  __shared__ int s_nbig;
  float4* part = reinterpret_cast<float4*>(gsm);
  constexpr int NG = 32 / G;
  constexpr int C4 = G * R;
  constexpr int kGroups = kGatherWarps * NG;
  constexpr int BS = (R == 1) ? (G < 8 ? G : 8) : (R == 2 ? 4 : 2);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane / G, gl = lane % G;
  const bool gact = g < NG;
  const int gid = warp * NG + g;
  const int gbase = g * G;
  const int CW = TX + 1, CH = TY + 1, ncell = CW * CH;
  float4* scratch = part + 4 * ncell * C4;
  int t = blockIdx.x;
  const int txi = t % tiles_x; t /= tiles_x;
  const int tyi = t % tiles_y;
  const int map = t / tiles_y;
  const int x0 = txi * TX, y0 = tyi * TY;
  const int64_t map_base = (int64_t)map * p.H * p.W;
  const float4* __restrict__ ghat4 = reinterpret_cast<const float4*>(p.ghat);
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (threadIdx.x == 0) s_nbig = 0;
  __syncthreads();

  // ---- phase 1: cells -> four corner partial sums each ------------------------------------------
  // (row, col) of this group's cell inside the (CH x CW) cell window, advanced by kGroups cells per round
  const int step_r = kGroups / CW, step_c = kGroups - step_r * CW;
  int crow = gid / CW, ccol = gid - crow * CW;
  for (int c0 = 0; c0 < ncell; c0 += kGroups) {
    const int cl = c0 + gid;
    const bool slot = gact && cl < ncell;
    const int cy = y0 - 1 + crow, cx = x0 - 1 + ccol;
    crow += step_r; ccol += step_c;
    if (ccol >= CW) { ccol -= CW; ++crow; }
    int s = 0, e = 0;
    if (slot) {
      if (cy >= 0 && cx >= 0 && cy < p.H && cx < p.W) {
        const int64_t cell = map_base + (int64_t)cy * p.W + cx;
        s = __ldg(p.bin_start + (cell << p.nb_log2));
        e = __ldg(p.bin_start + ((cell + 1) << p.nb_log2));
      }
      if (e - s >= kBigCell) {  // defer to the cooperative pass (list order is irrelevant: cells are independent)
        int idx = 0;
        if (gl == 0) idx = atomicAdd(&s_nbig, 1);
        idx = __shfl_sync(__activemask(), idx, gbase);
        if (idx < kMaxBig) {
          if (gl == 0) s_big[idx] = cl;
          e = s;
        }
      }
    }
    const int kmax = __reduce_max_sync(kFullB, e - s);
    float4 anw[R], ane[R], asw[R], ase[R];
#pragma unroll
    for (int i = 0; i < R; ++i) { anw[i] = zero4; ane[i] = zero4; asw[i] = zero4; ase[i] = zero4; }
    // lane j of the group fetches entry k0+j; the next batch's entries are fetched while the current rows are in flight
    int4 mine = make_int4(-1, 0, 0, 0);
    if (gl < BS && (s + gl) < e) mine = __ldg(p.sorted + s + gl);
    for (int k0 = 0; k0 < kmax; k0 += BS) {
      const int4 cur = mine;
      mine = make_int4(-1, 0, 0, 0);
      if (gl < BS && (s + k0 + BS + gl) < e) mine = __ldg(p.sorted + s + k0 + BS + gl);
      gather_batch<G, R, BS>(cur, gbase, gl, gact, ghat4, anw, ane, asw, ase);
    }
    if (slot) {
#pragma unroll
      for (int i = 0; i < R; ++i) {
        const int o = cl * C4 + i * G + gl;
        part[o] = anw[i];
        part[ncell * C4 + o] = ane[i];
        part[2 * ncell * C4 + o] = asw[i];
        part[3 * ncell * C4 + o] = ase[i];
      }
    }
  }
  __syncthreads();
  // ---- phase 1b: heavily populated cells, all lane groups of the CTA together -----------------------
  // Group q takes the batches q, q+kGroups, ... of the cell; the group sums are then added in the fixed order
  // g = 0..NG-1 inside each warp and w = 0..kGatherWarps-1 across warps (both through the scratch area) -- a pure function of
  // the cell's entry count, hence deterministic.
  const int nbig = min(s_nbig, kMaxBig);
  for (int bi = 0; bi < nbig; ++bi) {
    const int cl = s_big[bi];
    const int cy = y0 - 1 + cl / CW, cx = x0 - 1 + cl % CW;
    const int64_t cell = map_base + (int64_t)cy * p.W + cx;
    const int s = __ldg(p.bin_start + (cell << p.nb_log2)), e = __ldg(p.bin_start + ((cell + 1) << p.nb_log2));
    const int nbatch = (e - s + BS - 1) / BS;
    const int rounds = (nbatch + kGroups - 1) / kGroups;
    float4 anw[R], ane[R], asw[R], ase[R];
#pragma unroll
    for (int i = 0; i < R; ++i) { anw[i] = zero4; ane[i] = zero4; asw[i] = zero4; ase[i] = zero4; }
    for (int r = 0; r < rounds; ++r) {
      const int k0 = (r * kGroups + gid) * BS;
      int4 cur = make_int4(-1, 0, 0, 0);
      if (gact && gl < BS && (s + k0 + gl) < e) cur = __ldg(p.sorted + s + k0 + gl);
      gather_batch<G, R, BS>(cur, gbase, gl, gact, ghat4, anw, ane, asw, ase);
    }
    // inside the warp: the groups add their sums into the warp's scratch slot one after the other (fixed order)
#pragma unroll 1
    for (int gg = 0; gg < NG; ++gg) {
      if (gact && g == gg) {
#pragma unroll
        for (int i = 0; i < R; ++i) {
          const int o = i * G + gl;
          float4* q0 = scratch + (warp * 4 + 0) * C4 + o;
          float4* q1 = scratch + (warp * 4 + 1) * C4 + o;
          float4* q2 = scratch + (warp * 4 + 2) * C4 + o;
          float4* q3 = scratch + (warp * 4 + 3) * C4 + o;
          if (gg == 0) { *q0 = anw[i]; *q1 = ane[i]; *q2 = asw[i]; *q3 = ase[i]; }
          else { *q0 = add4(*q0, anw[i]); *q1 = add4(*q1, ane[i]); *q2 = add4(*q2, asw[i]); *q3 = add4(*q3, ase[i]); }
        }
      }
      __syncwarp();
    }
    __syncthreads();
    for (int o = threadIdx.x; o < 4 * C4; o += kGatherWarps * 32) {  // o = corner * C4 + quad
      float4 a = scratch[o];
#pragma unroll
      for (int w = 1; w < kGatherWarps; ++w) a = add4(a, scratch[w * 4 * C4 + o]);
      const int corner = o / C4, q = o - corner * C4;
      part[(corner * ncell + cl) * C4 + q] = a;
    }
    __syncthreads();
  }
  __syncthreads();
  // ---- phase 2: texel (y,x) = nw(y,x) + ne(y,x-1) + sw(y-1,x) + se(y-1,x-1), in that order ---------
  const int64_t hw = (int64_t)p.H * p.W;
  for (int q0 = 0; q0 < TX * TY; q0 += kGroups) {
    const int q = q0 + gid;
    if (!gact || q >= TX * TY) continue;
    const int ty = q / TX, tx = q - ty * TX;
    const int y = y0 + ty, x = x0 + tx;
    if (y >= p.H || x >= p.W) continue;
    const int cnw = (ty + 1) * CW + tx + 1, cne = cnw - 1, csw = cnw - CW, cse = csw - 1;
    const int64_t tex = map_base + (int64_t)y * p.W + x;
#pragma unroll
    for (int i = 0; i < R; ++i) {
      const int o = i * G + gl;
      float4 a = part[cnw * C4 + o];
      const float4 b1 = part[(ncell + cne) * C4 + o];
      const float4 b2 = part[(2 * ncell + csw) * C4 + o];
      const float4 b3 = part[(3 * ncell + cse) * C4 + o];
      a.x = __fadd_rn(__fadd_rn(__fadd_rn(a.x, b1.x), b2.x), b3.x);
      a.y = __fadd_rn(__fadd_rn(__fadd_rn(a.y, b1.y), b2.y), b3.y);
      a.z = __fadd_rn(__fadd_rn(__fadd_rn(a.z, b1.z), b2.z), b3.z);
      a.w = __fadd_rn(__fadd_rn(__fadd_rn(a.w, b1.w), b2.w), b3.w);
      if (p.xchg_world) {
        // owner's staging slot of this rank, channels-last: one 16-byte store per lane, C contiguous floats per texel
        const int v = map / p.B, b = map - v * p.B;
        const int owner = v / p.xchg_vpo, vl = v - owner * p.xchg_vpo;
        float4* dst = reinterpret_cast<float4*>(p.xchg_peer[owner]) +
                      ((((int64_t)p.xchg_rank * p.xchg_vpo + vl) * p.B + b) * hw + (int64_t)y * p.W + x) * C4;
        dst[o] = a;
      } else if (!p.grad_nchw) {
        reinterpret_cast<float4*>(p.grad_feats)[tex * C4 + o] = a;
      } else {
        // (V,B,C,H,W): channel c of texel (map, y, x) lives at ((map*C + c)*H + y)*W + x
        float* dst = p.grad_feats + (int64_t)map * p.C * hw + (int64_t)y * p.W + x + (int64_t)(o * 4) * hw;
        dst[0] = a.x; dst[hw] = a.y; dst[2 * hw] = a.z; dst[3 * hw] = a.w;
      }
    }
  }
  // the cells whose texel (cy, cx) lies in this tile are owned by it: hand their claim counters back as zeros, so that
  // the binning state can serve another backward call (BinLayout).  Last, off the critical path: plain stores.
  for (int r = 0; r < TY; ++r) {
    const int cy = y0 + r;
    if (cy >= p.H) break;
    const int cxb = min(x0 + TX, p.W);
    const int64_t b0 = (map_base + (int64_t)cy * p.W + x0) << p.nb_log2;
    const int nbins = (cxb - x0) << p.nb_log2;
    for (int k = threadIdx.x; k < nbins; k += kGatherWarps * 32) p.bin_cursor[b0 + k] = 0;
  }
}

// ---- order: every bin's entries into ascending voxel order ----------------------------------------
// One thread per entry, out of place: rank = number of entries of the same bin with a smaller voxel index (voxel
// indices are unique within a bin because a voxel projects into a view at most once), destination = bin start +
// rank.  The lanes of a warp mostly sit in the same bin, so the k reads of the rank loop are L1 broadcasts.
// This removes the only nondeterminism of the backward pass (the claim order of the fill atomics).
// (Ranking inside the gather kernel instead -- one launch fewer -- was measured and dropped: per tile only a few hundred
// entries keep 256 threads busy, the halo cells are ranked twice and the whole CTA waits for its longest bin:
// fragment level 2 45 -> 61 us, dense level 2 106 -> 165 us, large scene 3.2 -> 5.6 ms; profiles/r02b_bench.json.)
constexpr int kOrderThreads = 256;
__global__ void __launch_bounds__(kOrderThreads) bp_bwd_order_kernel(const BwdParams p) {
  pdl_enter();
  const int total = __ldg(p.bin_start + p.Mb);
  for (int i = blockIdx.x * kOrderThreads + threadIdx.x; i < total; i += gridDim.x * kOrderThreads) {
    const int4 e = __ldg(p.entries + i);
    const int ms = __ldg(p.bin_start + e.w), me = __ldg(p.bin_start + e.w + 1);
    int rank = 0;
    const int* keys = reinterpret_cast<const int*>(p.entries);
    int j = ms;
    for (; j + 4 <= me; j += 4) {
      const int a0 = __ldg(keys + 4 * j), a1 = __ldg(keys + 4 * j + 4), a2 = __ldg(keys + 4 * j + 8),
                a3 = __ldg(keys + 4 * j + 12);
      rank += (a0 < e.x) + (a1 < e.x) + (a2 < e.x) + (a3 < e.x);
    }
    for (; j < me; ++j) rank += (__ldg(keys + 4 * j) < e.x) ? 1 : 0;
    p.sorted[ms + rank] = e;
  }
}

// any C: one warp per texel, lanes stride over channels
__global__ void __launch_bounds__(kGatherWarps * 32) bp_bwd_gather_generic_kernel(const BwdParams p) {
  pdl_enter();
  const int lane = threadIdx.x & 31;
  const int C = p.C;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t t = warp_global; t < p.M; t += nwarps) {
    const int x = (int)(t % p.W), y = (int)((t / p.W) % p.H);
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
#pragma unroll
    for (int corner = 0; corner < 4; ++corner) {
      const int dx = corner & 1, dy = corner >> 1;
      if (x - dx < 0 || y - dy < 0) continue;
      const int64_t cell = t - dx - (int64_t)dy * p.W;
      const int s = __ldg(p.bin_start + (cell << p.nb_log2)), e = __ldg(p.bin_start + ((cell + 1) << p.nb_log2));
      for (int k = s; k < e; ++k) {
        const int4 en = __ldg(p.sorted + k);
        const float fx = __int_as_float(en.y), fy = __int_as_float(en.z);
        const float wx = dx ? fx : __fsub_rn(1.0f, fx);
        const float wy = dy ? fy : __fsub_rn(1.0f, fy);
        const float w = __fmul_rn(wx, wy);
        const float* row = p.ghat + (int64_t)en.x * C;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int c = lane + 32 * i;
          if (c < C) acc[i] = __fadd_rn(acc[i], __fmul_rn(w, __ldg(row + c)));
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane + 32 * i;
      if (c < C) {
        if (!p.grad_nchw) {
          p.grad_feats[t * C + c] = acc[i];
        } else {
          const int64_t hw = (int64_t)p.H * p.W;
          const int64_t vb = t / hw, yx = t - vb * hw;
          p.grad_feats[(vb * C + c) * hw + yx] = acc[i];
        }
      }
    }
  }
}

// ---- host ----------------------------------------------------------------------------------------
struct BwdWs {
  size_t ghat, cnt, state, entries, sorted, total;
  BinLayout bl;
};

static BwdWs bwd_ws_layout(int64_t N, int B, int V, int C, int H, int W) {
  BwdWs w;
  w.bl = bin_layout(N, B, V, H, W);
  size_t o = 0;
  const size_t n1 = (size_t)(N > 0 ? N : 1);
  w.ghat = o; o = align_up(o + sizeof(float) * n1 * (size_t)C, 256);
  w.cnt = o; o = align_up(o + sizeof(float) * n1, 256);
  w.state = o; o = align_up(o + sizeof(int) * w.bl.total, 256);   // binning state when forward did not hand one over
  w.entries = o; o = align_up(o + sizeof(int4) * n1 * (size_t)V, 256);
  w.sorted = o; o = align_up(o + sizeof(int4) * n1 * (size_t)V, 256);
  w.total = o;
  return w;
}

typedef void (*gather_kernel_t)(const BwdParams, int, int, int, int);
static gather_kernel_t pick_gather_kernel(int C) {
  if (C % 4 != 0) return nullptr;
  const int q = C / 4;
#define D3M_BWD_CASE(g, r) \
  if (q == (g) * (r)) return bp_bwd_gather_tile_kernel<g, r>;
  D3M_BWD_CASE(6, 1)
  D3M_BWD_CASE(10, 1)
  D3M_BWD_CASE(10, 2)
  D3M_BWD_CASE(4, 1)
  D3M_BWD_CASE(8, 1)
  D3M_BWD_CASE(16, 1)
  D3M_BWD_CASE(8, 3)
  D3M_BWD_CASE(16, 2)
  D3M_BWD_CASE(2, 1)
  D3M_BWD_CASE(3, 1)
  D3M_BWD_CASE(5, 1)
#undef D3M_BWD_CASE
  return nullptr;
}

// texel tile of the gather kernel: the largest of a fixed ladder whose 4 x (TX+1)(TY+1) x C floats fit the budget
// Measured on B200 (profiles/r01_bp_gather_tile.txt): small tiles win -- more CTAs per SM hide the dependent
// bin_start -> entry -> row load chain better than a smaller halo helps -- down to 8x4 texels; C = 24 -> 8x8 (31 KB),
// C = 40 -> 8x4 (29 KB), C = 80 -> 8x4 (58 KB).
// Exception, measured in round 2 (profiles/r02z_gather_tile_density.txt): at C = 24 and FEW voxels per pixel (sparse
// fragment levels: ~6 voxels per pixel, ~4 entries per bilinear cell) the walk over cells, not the rows, is the cost, and
// the 16x8 tile (59 KB, 1.13 cells per texel instead of 1.27) is faster -- level-2 gather 49.9 -> 46.2 us, 64 fragments
// 10.9 -> 10.6 ms -- while dense volumes (46 voxels per pixel and more) stay on 8x8 (dense gather 115 vs 120 us).
static void pick_gather_tile(int C, double voxels_per_pixel, int& TX, int& TY, size_t& smem) {
  static const int ladder[][2] = {{16, 8}, {8, 8}, {8, 4}, {4, 4}, {2, 2}, {1, 1}};
  static const int env_kb = getenv("D3M_GATHER_SMEM_KB") ? atoi(getenv("D3M_GATHER_SMEM_KB")) : 0;  // tuning aid
  const bool sparse_small_c = C <= 24 && voxels_per_pixel < 16.0;
  for (int pass = 0; pass < 2; ++pass) {
    const size_t budget = env_kb ? (size_t)env_kb * 1024 : ((pass == 0 && !sparse_small_c) ? 32 * 1024 : 64 * 1024);
    for (auto& t : ladder) {
      TX = t[0]; TY = t[1];
      smem = (size_t)16 * C * (TX + 1) * (TY + 1);
      if (smem <= budget) break;
    }
    if (env_kb || TX * TY >= 32) break;  // never below 8x4 texels if 64 KB can hold it
  }
}

// `own_state`: the binning state lives in this call's workspace (no forward histogram): clear, histogram and scan it here.
template <int KIND>
static int launch_bwd(BwdParams p, const BinState& bins, const BinLayout& bl, bool own_state, float* cnt_ws,
                      cudaStream_t stream) {
  int sms = current_device_sms();
  const unsigned vox_ctas = (unsigned)((p.N + kSampleThreads - 1) / kSampleThreads);
  const unsigned vgroups = (unsigned)((p.V + kViewsPerThread - 1) / kViewsPerThread);
  if (cnt_ws) {  // no forward count handed over: recompute it
    LaunchScope ls("bp_bwd_count", stream);
    launch_k(bp_bwd_count_kernel<KIND>, dim3(vox_ctas), dim3(kSampleThreads), 0, stream, p, cnt_ws);
    D3M_CUDA_CHECK(cudaGetLastError());
  }
  if (own_state) {
    const int zrc = zero_async(bins.cnt, sizeof(int) * bl.zero_elems, stream);
    if (zrc != D3M_OK) return zrc;
    {
      LaunchScope ls("bp_bwd_hist", stream);
      launch_k(bp_bwd_hist_kernel<KIND>, dim3(vox_ctas, vgroups), dim3(kSampleThreads), 0, stream, p);
      D3M_CUDA_CHECK(cudaGetLastError());
    }
    ScanOnly so;
    so.bins = bins;
    LaunchScope ls("bp_bwd_scan", stream);
    launch_k(bp_bwd_scan_kernel, dim3((unsigned)bl.nchunks), dim3(kScanThreads), 0, stream, so);
    D3M_CUDA_CHECK(cudaGetLastError());
  }
  {
    const int rows_per_step = ((p.C & 3) == 0 && p.C <= 128) ? 32 / (p.C >> 2) : 1;
    const int64_t rows_per_cta = (int64_t)(kSampleThreads / 32) * kGhatSteps * rows_per_step;
    const int64_t ghat_blocks = (p.N + rows_per_cta - 1) / rows_per_cta;
    const int64_t fill_ctas = (int64_t)vox_ctas * vgroups;
    D3M_REQUIRE(fill_ctas + ghat_blocks < (1ll << 31), D3M_ERR_ARG, "back_project backward: too many fill CTAs");
    p.fill_ctas_x = (int)vox_ctas;
    p.ghat_ctas = (int)ghat_blocks;
    LaunchScope ls("bp_bwd_fill_ghat", stream);
    launch_k(bp_bwd_fill_ghat_kernel<KIND>, dim3((unsigned)(fill_ctas + ghat_blocks)), dim3(kSampleThreads), 0, stream, p);
    D3M_CUDA_CHECK(cudaGetLastError());
  }
  {
    int64_t ctas = (p.N * p.V + kOrderThreads - 1) / kOrderThreads;  // upper bound of the entry count (known on device only)
    if (ctas > (int64_t)sms * 16) ctas = (int64_t)sms * 16;
    if (ctas < 1) ctas = 1;
    LaunchScope ls("bp_bwd_order", stream);
    launch_k(bp_bwd_order_kernel, dim3((unsigned)ctas), dim3(kOrderThreads), 0, stream, p);
    D3M_CUDA_CHECK(cudaGetLastError());
  }
  gather_kernel_t k = pick_gather_kernel(p.C);
  if (k) {
    int TX, TY;
    size_t smem;
    pick_gather_tile(p.C, (double)p.N / ((double)p.B * p.H * p.W), TX, TY, smem);
    const int tiles_x = (p.W + TX - 1) / TX, tiles_y = (p.H + TY - 1) / TY;
    const int64_t tiles = (int64_t)p.V * p.B * tiles_x * tiles_y;
    D3M_REQUIRE(tiles < (1ll << 31), D3M_ERR_ARG, "back_project backward: too many gather tiles");
    smem += (size_t)kGatherWarps * 4 * p.C * 4;  // scratch of the cooperative big-cell pass
    D3M_CUDA_CHECK(ensure_dynamic_smem(reinterpret_cast<const void*>(k), smem));
    LaunchScope ls("bp_bwd_gather", stream);
    launch_k(k, dim3((unsigned)tiles), dim3(kGatherWarps * 32), smem, stream, p, TX, TY, tiles_x, tiles_y);
    D3M_CUDA_CHECK(cudaGetLastError());
  } else {
    D3M_REQUIRE(p.C <= 256, D3M_ERR_ARG, "back_project backward: C=%d unsupported (C%%4!=0 needs C<=256)", p.C);
    int64_t ctas = (p.M + kGatherWarps - 1) / kGatherWarps;
    if (ctas > (int64_t)sms * 16) ctas = (int64_t)sms * 16;
    if (ctas < 1) ctas = 1;
    {
      LaunchScope ls("bp_bwd_gather", stream);
      launch_k(bp_bwd_gather_generic_kernel, dim3((unsigned)ctas), dim3(kGatherWarps * 32), 0, stream, p);
      D3M_CUDA_CHECK(cudaGetLastError());
    }
    const int zrc = zero_async(bins.cursor, sizeof(int) * (size_t)bl.Mb, stream);  // the tile kernel does this itself
    if (zrc != D3M_OK) return zrc;
  }
  return D3M_OK;
}

}  // namespace d3m

using namespace d3m;

extern "C" size_t d3m_back_project_bwd_workspace(int64_t N, int B, int V, int C, int H, int W) {
  if (N < 0 || B < 1 || V < 1 || C < 1 || H < 1 || W < 1) return 0;
  return bwd_ws_layout(N, B, V, C, H, W).total;
}

struct Exchange {
  void* const* peer_bufs;   // host array of `world` device pointers (every rank's staging buffer, mapped here)
  int world, rank;
};

static int bwd_impl(const void* coords, int coords_kind, int64_t N, const float* origin, int B, float voxel_size, int V,
                    int C, int H, int W, const float* KRcam, const float* grad_out, const float* count, int* cell_hist,
                    float* grad_feats_nhwc, int grad_nchw, void* workspace, size_t workspace_bytes, const Exchange* xc,
                    void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  D3M_REQUIRE(d3m_device_count() > 0, D3M_ERR_NO_DEVICE,
              "back_project backward: no CUDA device (there is no CPU fallback)");
  D3M_REQUIRE(N >= 0 && B >= 1 && V >= 1 && C >= 1 && H >= 2 && W >= 2, D3M_ERR_ARG,
              "back_project backward: bad sizes N=%lld B=%d V=%d C=%d H=%d W=%d", (long long)N, B, V, C, H, W);
  D3M_REQUIRE(coords_kind >= 0 && coords_kind <= 2, D3M_ERR_ARG, "back_project backward: coords_kind=%d", coords_kind);
  D3M_REQUIRE((int64_t)V * B * H * W < (1ll << 30), D3M_ERR_ARG, "back_project backward: V*B*H*W must be < 2^30");
  D3M_REQUIRE(V <= 65535, D3M_ERR_ARG, "back_project backward: too many views");
  D3M_REQUIRE(N * (int64_t)V < (1ll << 31), D3M_ERR_ARG, "back_project backward: N*V must be < 2^31 samples");
  D3M_REQUIRE((grad_feats_nhwc || xc) && workspace, D3M_ERR_ARG, "back_project backward: NULL pointer");
  const BwdWs w = bwd_ws_layout(N, B, V, C, H, W);
  int vpo = 0;
  if (xc) {
    D3M_REQUIRE(xc->peer_bufs && xc->world >= 1 && xc->world <= kMaxExchangeRanks && xc->rank >= 0 && xc->rank < xc->world,
                D3M_ERR_ARG, "back_project backward exchange: bad world / rank (at most %d ranks)", kMaxExchangeRanks);
    D3M_REQUIRE((C & 3) == 0 && pick_gather_kernel(C) != nullptr, D3M_ERR_ARG,
                "back_project backward exchange: C=%d has no tiled gather kernel", C);
    vpo = (V + xc->world - 1) / xc->world;
    for (int r = 0; r < xc->world; ++r)
      D3M_REQUIRE(xc->peer_bufs[r] && aligned16(xc->peer_bufs[r]), D3M_ERR_ALIGN,
                  "back_project backward exchange: staging buffer of rank %d is NULL or unaligned", r);
  }
  if (N == 0) {
    if (!xc) {
      D3M_CUDA_CHECK(cudaMemsetAsync(grad_feats_nhwc, 0, sizeof(float) * (size_t)w.bl.M * C, stream));
      return D3M_OK;
    }
    // an empty shard still owes every owner a slot of zeros
    const size_t slot = sizeof(float) * (size_t)vpo * B * H * W * C;
    for (int r = 0; r < xc->world; ++r)
      D3M_CUDA_CHECK(cudaMemsetAsync(static_cast<unsigned char*>(xc->peer_bufs[r]) + slot * xc->rank, 0, slot, stream));
    return D3M_OK;
  }
  D3M_REQUIRE(coords && origin && KRcam && grad_out, D3M_ERR_ARG, "back_project backward: NULL pointer");
  D3M_REQUIRE(aligned16(coords) && aligned16(KRcam) && aligned16(grad_feats_nhwc) && aligned16(workspace) &&
                  aligned16(cell_hist),
              D3M_ERR_ALIGN, "back_project backward: coords/KRcam/grad_feats/cell_hist/workspace must be 16-byte aligned");
  D3M_REQUIRE(workspace_bytes >= w.total, D3M_ERR_WORKSPACE, "back_project backward: workspace %zu < %zu",
              workspace_bytes, w.total);
  PdlScope pdl_scope(stream, (long long)N * V);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  const bool own_state = cell_hist == nullptr;
  const BinState bins = bin_state(own_state ? reinterpret_cast<int*>(ws + w.state) : cell_hist, w.bl);
  BwdParams p;
  memset(&p, 0, sizeof(p));
  p.coords = coords; p.N = N; p.origin = origin; p.B = B; p.vs = voxel_size;
  p.V = V; p.C = C; p.H = H; p.W = W; p.KR = KRcam; p.grad_out = grad_out;
  p.ghat = reinterpret_cast<float*>(ws + w.ghat);
  p.bin_cnt = bins.cnt;
  p.bin_cursor = bins.cursor;
  p.bin_start = bins.start;
  p.entries = reinterpret_cast<int4*>(ws + w.entries);
  p.sorted = reinterpret_cast<int4*>(ws + w.sorted);
  p.grad_feats = grad_feats_nhwc;
  p.M = w.bl.M; p.Mb = w.bl.Mb; p.nb_log2 = w.bl.nb_log2;
  p.grad_nchw = grad_nchw ? 1 : 0;
  if (xc) {
    p.xchg_world = xc->world; p.xchg_rank = xc->rank; p.xchg_vpo = vpo;
    for (int r = 0; r < xc->world; ++r) p.xchg_peer[r] = static_cast<float*>(xc->peer_bufs[r]);
  }
  float* cnt_ws = count ? nullptr : reinterpret_cast<float*>(ws + w.cnt);
  p.count = count ? count : cnt_ws;
  if (coords_kind == D3M_COORDS_F32) return launch_bwd<D3M_COORDS_F32>(p, bins, w.bl, own_state, cnt_ws, stream);
  if (coords_kind == D3M_COORDS_I64) return launch_bwd<D3M_COORDS_I64>(p, bins, w.bl, own_state, cnt_ws, stream);
  return launch_bwd<D3M_COORDS_I32>(p, bins, w.bl, own_state, cnt_ws, stream);
}

extern "C" int d3m_back_project_bwd(const void* coords, int coords_kind, int64_t N, const float* origin, int B,
                                    float voxel_size, int V, int C, int H, int W, const float* KRcam,
                                    const float* grad_out, const float* count, int* cell_hist,
                                    float* grad_feats_nhwc, int grad_nchw, void* workspace, size_t workspace_bytes,
                                    void* stream_) {
  return bwd_impl(coords, coords_kind, N, origin, B, voxel_size, V, C, H, W, KRcam, grad_out, count, cell_hist,
                  grad_feats_nhwc, grad_nchw, workspace, workspace_bytes, nullptr, stream_);
}

extern "C" int d3m_back_project_bwd_exchange(const void* coords, int coords_kind, int64_t N, const float* origin, int B,
                                             float voxel_size, int V, int C, int H, int W, const float* KRcam,
                                             const float* grad_out, const float* count, int* cell_hist,
                                             void* const* peer_staging_host, int world, int rank, void* workspace,
                                             size_t workspace_bytes, void* stream_) {
  Exchange xc;
  xc.peer_bufs = peer_staging_host; xc.world = world; xc.rank = rank;
  return bwd_impl(coords, coords_kind, N, origin, B, voxel_size, V, C, H, W, KRcam, grad_out, count, cell_hist, nullptr, 0,
                  workspace, workspace_bytes, &xc, stream_);
}

// ---- owner side of the exchange: sum the `world` slots of this rank's staging buffer in rank order ----------------------
// staging (world, vpo, B, H, W, C) channels-last -> out (n_views, B, C, H, W) in the reference layout; n_views <= vpo is the
// number of views this rank really owns.  One thread per (texel, channel quad); the slots are added in ascending rank
// order, so the result does not depend on arrival order.
namespace d3m {
__global__ void __launch_bounds__(256) grad_slots_sum_kernel(const float4* __restrict__ staging, float* __restrict__ out,
                                                             int world, int vpo, int n_views, int B, int C4, int64_t hw) {
  pdl_enter();
  const int64_t total = (int64_t)n_views * B * hw * C4;
  const int64_t slot = (int64_t)vpo * B * hw * C4;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    float4 a = __ldcs(staging + i);
    for (int r = 1; r < world; ++r) {
      const float4 t = __ldcs(staging + r * slot + i);
      a.x = __fadd_rn(a.x, t.x); a.y = __fadd_rn(a.y, t.y); a.z = __fadd_rn(a.z, t.z); a.w = __fadd_rn(a.w, t.w);
    }
    const int q = (int)(i % C4);
    const int64_t tex = i / C4;              // (view, b, y, x) linear
    const int64_t yx = tex % hw, vb = tex / hw;
    float* dst = out + (vb * C4 * 4 + (int64_t)q * 4) * hw + yx;
    dst[0] = a.x; dst[hw] = a.y; dst[2 * hw] = a.z; dst[3 * hw] = a.w;
  }
}
}  // namespace d3m

extern "C" int d3m_grad_slots_sum(const float* staging, int world, int views_per_owner, int n_views, int B, int C, int H,
                                  int W, float* out_nchw, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  D3M_REQUIRE(d3m_device_count() > 0, D3M_ERR_NO_DEVICE, "grad_slots_sum: no CUDA device");
  D3M_REQUIRE(world >= 1 && views_per_owner >= 1 && n_views >= 0 && n_views <= views_per_owner && B >= 1 && C >= 4 &&
                  (C & 3) == 0 && H >= 1 && W >= 1,
              D3M_ERR_ARG, "grad_slots_sum: bad arguments");
  if (n_views == 0) return D3M_OK;
  D3M_REQUIRE(staging && out_nchw && aligned16(staging), D3M_ERR_ARG, "grad_slots_sum: NULL / unaligned pointer");
  const int64_t total = (int64_t)n_views * B * H * W * (C / 4);
  int64_t ctas = (total + 255) / 256;
  if (ctas > 148 * 16) ctas = 148 * 16;
  LaunchScope ls("grad_slots_sum", stream);
  launch_k(grad_slots_sum_kernel, dim3((unsigned)ctas), dim3(256), 0, stream, reinterpret_cast<const float4*>(staging),
           out_nchw, world, views_per_owner, n_views, B, C / 4, (int64_t)H * W);
  D3M_CUDA_CHECK(cudaGetLastError());
  return D3M_OK;
}

