// back_project backward (grad w.r.t. the 2D feature maps) for sm_100a -- deterministic, no float atomics.
// Replaces autograd through deep3dmap/core/voxel/back_project.py:55-73 of the reference:
//   grad_sample[v,n,:] = grad_out[n,:C] / max(count[n],1)   masked to valid views,
//   grid_sampler_2d_backward: grad_feats[v,b,y,x,:] += w_corner(v,n) * grad_sample[v,n,:]  (aten uses atomicAdd).
//
// The scatter is turned into a gather (texel-centric), so every texel is produced by exactly one lane group
// in a fixed order:
//   prep    per voxel: project into every view (same arithmetic as forward), count valid views, histogram the
//           valid samples per bilinear cell (v,b,y0,x0) with INTEGER atomics, and write the pre-divided rows
//           ghat[n,:] = grad_out[n,:C]/max(count,1) into a 16-byte aligned (N,C) buffer (grad_out rows are
//           (C+1) floats and cannot be vector-loaded).
//   scan    exclusive prefix sum of the cell histogram (two kernels, ticket-finalised chunk sums).
//   fill    project again, claim a slot in the cell with an integer atomic, store {n, fx, fy} (16 B).
//   order   sort every cell's entries by voxel index (windowed warp rank-sort; in-place bitonic for cells
//           with more than 32 entries) -- this removes the only nondeterminism (slot claim order).
//   gather  lane group per texel: walk the 4 neighbouring cells (as nw, ne, sw, se corner), per entry one
//           128-bit broadcast load of the entry and R 128-bit loads of the ghat row per lane; plain fp32
//           mul + add in entry order; one coalesced 128-bit store of the texel's gradient.
#include "d3m_common.cuh"

namespace d3m {

constexpr unsigned kFullB = 0xffffffffu;
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanChunk = kScanThreads * kScanItems;
constexpr int kGatherWarps = 8;

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
  int* bin_cnt;
  int* bin_start;  // M+1
  int4* entries;
  float* grad_feats;
  int64_t M;  // V*B*H*W cells
  unsigned long long* scan_state;  // nchunks words, zeroed together with bin_cnt
  unsigned int* counter;           // scan ticket, zeroed together with bin_cnt
  int nchunks;
  int grad_nchw;                   // 1: gather writes (V,B,C,H,W) directly
};

constexpr int kSampleThreads = 256;
constexpr int kGhatPerThread = 4;

// Only used when the caller does not hand over the forward pass's `count`: valid views per voxel.
template <int KIND>
__global__ void __launch_bounds__(kSampleThreads) bp_bwd_count_kernel(const BwdParams p, float* __restrict__ cnt_out) {
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

// One thread per (voxel, view) pair: blockIdx.y = view for y < V.  Short dependency chains and N*V-way
// parallelism instead of a 9-deep serial loop per voxel (these passes are latency-bound at fragment size).
//   FILL = false: histogram the valid samples per bilinear cell (integer RED); the extra grid rows y >= V
//                 compute ghat[n,c] = grad_out[n,c] / max(count[n],1) (div backward of back_project.py:72),
//                 re-packed to 16-byte aligned rows of C floats, one element per thread-iteration.
//   FILL = true : claim a slot in the cell (integer atomic), store {n, fx, fy}.
template <int KIND, bool FILL>
__global__ void __launch_bounds__(kSampleThreads) bp_bwd_sample_kernel(const BwdParams p) {
  const int v = blockIdx.y;
  if (!FILL && v >= p.V) {
    const int C = p.C, C1 = C + 1;
    const int64_t total = p.N * C;
    const int64_t blk = (int64_t)(v - p.V) * gridDim.x + blockIdx.x;
    const int64_t base = blk * (kSampleThreads * kGhatPerThread) + threadIdx.x;
    float g[kGhatPerThread], d[kGhatPerThread];
#pragma unroll
    for (int k = 0; k < kGhatPerThread; ++k) {  // all loads first
      const int64_t i = base + (int64_t)k * kSampleThreads;
      g[k] = 0.f; d[k] = 1.f;
      if (i < total) {
        const int64_t r = i / C;
        const int c = (int)(i - r * C);
        g[k] = __ldg(p.grad_out + r * C1 + c);
        d[k] = fmaxf(__ldg(p.count + r), 1.0f);
      }
    }
#pragma unroll
    for (int k = 0; k < kGhatPerThread; ++k) {
      const int64_t i = base + (int64_t)k * kSampleThreads;
      if (i < total) p.ghat[i] = __fdiv_rn(g[k], d[k]);
    }
    return;
  }
  const int64_t n = (int64_t)blockIdx.x * kSampleThreads + threadIdx.x;
  if (n >= p.N) return;
  float cx, cy, cz;
  const int b = load_coord<KIND>(p.coords, n, p.B, cx, cy, cz);
  if (b < 0) return;
  float gx, gy, gz;
  const float* o = p.origin + 3 * b;
  voxel_world(cx, cy, cz, p.vs, __ldg(o), __ldg(o + 1), __ldg(o + 2), gx, gy, gz);
  float4 r0, r1, r2;
  load_krcam(p.KR, v, p.B, b, r0, r1, r2);
  const Sample s = project(gx, gy, gz, r0, r1, r2, (float)(p.W - 1), (float)(p.H - 1));
  if (!s.valid) return;
  const int key = ((v * p.B + b) * p.H + s.y0) * p.W + s.x0;
  if (!FILL) {
    atomicAdd(p.bin_cnt + key, 1);
  } else {
    const int slot = atomicSub(p.bin_cnt + key, 1) - 1;  // counts back to zero; order fixed later by `order`
    const int pos = __ldg(p.bin_start + key) + slot;
    p.entries[pos] = make_int4((int)n, __float_as_int(s.fx), __float_as_int(s.fy), 0);
  }
}

// ---- exclusive scan of the cell histogram: ONE pass, chained look-back ---------------------------
// CTAs take chunk ids from a ticket (so a chunk's predecessors are always already scheduled), publish their
// aggregate, walk back over predecessors until they meet an inclusive prefix, then publish their own.
// state word = (flag << 32) | value, flag 0 = not ready, 1 = aggregate, 2 = inclusive prefix.  Integer sums:
// the result does not depend on the order in which CTAs arrive.
__global__ void __launch_bounds__(kScanThreads) bp_scan_kernel(const BwdParams p) {
  __shared__ int red[kScanThreads / 32];
  __shared__ int s_cid, s_prefix;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_cid = (int)atomicAdd(p.counter, 1u);
  __syncthreads();
  const int cid = s_cid;
  const int64_t base = (int64_t)cid * kScanChunk + (int64_t)tid * kScanItems;
  int v[kScanItems];
  if (base + kScanItems <= p.M) {
    const int4* q = reinterpret_cast<const int4*>(p.bin_cnt + base);
#pragma unroll
    for (int i = 0; i < kScanItems / 4; ++i) {
      const int4 t = q[i];
      v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) v[i] = (base + i < p.M) ? p.bin_cnt[base + i] : 0;
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) s += v[i];
  int inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(kFullB, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) red[warp] = inc;
  __syncthreads();
  int woff = 0, total = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w) {
    if (w < warp) woff += red[w];
    total += red[w];
  }
  if (tid == 0) {
    volatile unsigned long long* st = p.scan_state;
    int prefix = 0;
    if (cid == 0) {
      st[0] = (2ull << 32) | (unsigned)total;
    } else {
      st[cid] = (1ull << 32) | (unsigned)total;
      __threadfence();
      int j = cid - 1;
      while (true) {
        unsigned long long w;
        do { w = st[j]; } while ((w >> 32) == 0ull);
        prefix += (int)(unsigned)(w & 0xffffffffull);
        if ((w >> 32) == 2ull) break;
        --j;
      }
      st[cid] = (2ull << 32) | (unsigned)(prefix + total);
    }
    s_prefix = prefix;
    if (cid == gridDim.x - 1) p.bin_start[p.M] = prefix + total;
  }
  __syncthreads();
  int off = s_prefix + woff + inc - s;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < p.M) p.bin_start[base + i] = off;
    off += v[i];
  }
}

// ---- order: sort each cell's entries by voxel index ---------------------------------------------
__device__ __forceinline__ void bitonic_cell(volatile int4* E, int k, int lane) {
  // all comparators ascending (mirror first stage), so virtual +inf padding above k never moves
  int K2 = 1;
  while (K2 < k) K2 <<= 1;
  for (int sz = 2; sz <= K2; sz <<= 1) {
    const int half = sz >> 1;
    for (int i = lane; i < (K2 >> 1); i += 32) {
      const int blk = i / half, off = i - blk * half;
      const int lo = blk * sz + off, hi = blk * sz + sz - 1 - off;
      if (hi < k) {
        const int a = E[lo].x, b = E[hi].x;
        if (a > b) {
          const int4 ea = make_int4(E[lo].x, E[lo].y, E[lo].z, E[lo].w);
          const int4 eb = make_int4(E[hi].x, E[hi].y, E[hi].z, E[hi].w);
          E[lo].x = eb.x; E[lo].y = eb.y; E[lo].z = eb.z; E[lo].w = eb.w;
          E[hi].x = ea.x; E[hi].y = ea.y; E[hi].z = ea.z; E[hi].w = ea.w;
        }
      }
    }
    __syncwarp();
    for (int st = half >> 1; st >= 1; st >>= 1) {
      for (int i = lane; i < (K2 >> 1); i += 32) {
        const int blk = i / st, off = i - blk * st;
        const int lo = blk * 2 * st + off, hi = lo + st;
        if (hi < k) {
          const int a = E[lo].x, b = E[hi].x;
          if (a > b) {
            const int4 ea = make_int4(E[lo].x, E[lo].y, E[lo].z, E[lo].w);
            const int4 eb = make_int4(E[hi].x, E[hi].y, E[hi].z, E[hi].w);
            E[lo].x = eb.x; E[lo].y = eb.y; E[lo].z = eb.z; E[lo].w = eb.w;
            E[hi].x = ea.x; E[hi].y = ea.y; E[hi].z = ea.z; E[hi].w = ea.w;
          }
        }
      }
      __syncwarp();
    }
  }
}


constexpr int kOrderWindow = 256;  // entries staged per warp (4 KB of shared memory)

// Each warp owns a run of cells.  Per iteration it stages a window of whole cells (<= 32 cells, <= 256 entries)
// in shared memory, ranks every entry inside its own cell by counting smaller voxel indices (voxel indices are
// unique within a cell), and writes the window back in ascending order.  Cells larger than the window fall
// back to an in-place bitonic sort in global memory.
__global__ void __launch_bounds__(256) bp_bwd_order_kernel(const BwdParams p, const int bins_per_task) {
  __shared__ int4 s_ent[8][kOrderWindow];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  int4* win = s_ent[wib];
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t ntasks = (p.M + bins_per_task - 1) / bins_per_task;
  for (int64_t task = warp_global; task < ntasks; task += nwarps) {
    int64_t bin = task * bins_per_task;
    const int64_t bin_hi = min(p.M, bin + bins_per_task);
    while (bin < bin_hi) {
      const int64_t mb = bin + lane;  // lane l looks at cell bin+l
      const bool have = mb < bin_hi;
      const int sb = have ? __ldg(p.bin_start + mb) : 0x7fffffff;
      const int eb = have ? __ldg(p.bin_start + mb + 1) : 0x7fffffff;
      const int s0 = __shfl_sync(kFullB, sb, 0);
      // whole cells that fit into the window starting at s0 (eb is monotone -> leading run of trues)
      const unsigned fits = __ballot_sync(kFullB, have && (eb - s0) <= kOrderWindow);
      const int nb = __popc(fits);
      if (nb == 0) {  // first cell alone exceeds the window
        const int e0 = __shfl_sync(kFullB, eb, 0);
        bitonic_cell(reinterpret_cast<volatile int4*>(p.entries + s0), e0 - s0, lane);
        bin += 1;
        continue;
      }
      const int e_end = __shfl_sync(kFullB, eb, nb - 1);
      const int total = e_end - s0;
      const unsigned need = __ballot_sync(kFullB, lane < nb && (eb - sb) >= 2);
      if (need != 0u) {
        for (int i = lane; i < total; i += 32) win[i] = p.entries[s0 + i];
        __syncwarp();
        for (int i0 = 0; i0 < total; i0 += 32) {
          const int i = i0 + lane;
          const bool on = i < total;
          const int pos = s0 + i;
          // my cell: the last of the nb cells whose start is <= pos (empty cells share their successor's start)
          int ms = s0, me = s0;
          for (int l = 0; l < nb; ++l) {
            const int vs_ = __shfl_sync(kFullB, sb, l), ve = __shfl_sync(kFullB, eb, l);
            if (on && vs_ <= pos && pos < ve) { ms = vs_; me = ve; }
          }
          if (on && me - ms >= 2) {
            const int4 e = win[i];
            int rank = 0;
            for (int j = ms - s0; j < me - s0; ++j) rank += (win[j].x < e.x) ? 1 : 0;
            p.entries[ms + rank] = e;
          }
        }
        __syncwarp();
      }
      bin += nb;
    }
  }
}

// ---- gather --------------------------------------------------------------------------------------
template <int G, int R>
__global__ void __launch_bounds__(kGatherWarps * 32) bp_bwd_gather_kernel(const BwdParams p) {
  constexpr int NG = 32 / G;
  const int lane = threadIdx.x & 31;
  const int g = lane / G, gl = lane % G;
  const int C4 = p.C >> 2;
  const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float4* __restrict__ ghat4 = reinterpret_cast<const float4*>(p.ghat);
  float4* __restrict__ grad4 = reinterpret_cast<float4*>(p.grad_feats);
  const int64_t nsteps = (p.M + NG - 1) / NG;
  for (int64_t step = warp_global; step < nsteps; step += nwarps) {
    const int64_t t = step * NG + g;
    const bool gvalid = (g < NG) && (t < p.M);
    const int x = gvalid ? (int)(t % p.W) : 0;
    const int y = gvalid ? (int)((t / p.W) % p.H) : 0;
    float4 acc[R];
#pragma unroll
    for (int i = 0; i < R; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int corner = 0; corner < 4; ++corner) {
      // corner 0: this texel is the nw corner of cell (y,x); 1: ne of (y,x-1); 2: sw of (y-1,x); 3: se of (y-1,x-1)
      const int dx = corner & 1, dy = corner >> 1;
      const bool exists = gvalid && (x - dx >= 0) && (y - dy >= 0);
      const int64_t cell = t - dx - (int64_t)dy * p.W;
      int s = 0, e = 0;
      if (exists) { s = __ldg(p.bin_start + cell); e = __ldg(p.bin_start + cell + 1); }
      const int kmax = __reduce_max_sync(kFullB, e - s);
      for (int k = 0; k < kmax; ++k) {
        const bool on = (s + k) < e;
        int4 en = make_int4(0, 0, 0, 0);
        if (on) en = __ldg(p.entries + s + k);
        const float fx = __int_as_float(en.y), fy = __int_as_float(en.z);
        const float wx = dx ? fx : __fsub_rn(1.0f, fx);
        const float wy = dy ? fy : __fsub_rn(1.0f, fy);
        const float w = __fmul_rn(wx, wy);
        const float4* row = ghat4 + (int64_t)en.x * C4 + gl;
#pragma unroll
        for (int i = 0; i < R; ++i) {
          if (on) {
            const float4 gq = __ldg(row + i * G);
            acc[i].x = __fadd_rn(acc[i].x, __fmul_rn(w, gq.x));
            acc[i].y = __fadd_rn(acc[i].y, __fmul_rn(w, gq.y));
            acc[i].z = __fadd_rn(acc[i].z, __fmul_rn(w, gq.z));
            acc[i].w = __fadd_rn(acc[i].w, __fmul_rn(w, gq.w));
          }
        }
      }
    }
    if (gvalid) {
      if (!p.grad_nchw) {
#pragma unroll
        for (int i = 0; i < R; ++i) grad4[t * C4 + i * G + gl] = acc[i];
      } else {
        // (V,B,C,H,W): texel t = (vb, y, x) -> channel c lives at ((vb*C + c)*H + y)*W + x
        const int64_t hw = (int64_t)p.H * p.W;
        const int64_t vb = t / hw, yx = t - vb * hw;
        float* dst = p.grad_feats + vb * p.C * hw + yx;
#pragma unroll
        for (int i = 0; i < R; ++i) {
          const int c0 = (i * G + gl) * 4;
          dst[(int64_t)(c0 + 0) * hw] = acc[i].x;
          dst[(int64_t)(c0 + 1) * hw] = acc[i].y;
          dst[(int64_t)(c0 + 2) * hw] = acc[i].z;
          dst[(int64_t)(c0 + 3) * hw] = acc[i].w;
        }
      }
    }
  }
}

// any C: one warp per texel, lanes stride over channels
__global__ void __launch_bounds__(kGatherWarps * 32) bp_bwd_gather_generic_kernel(const BwdParams p) {
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
      const int s = __ldg(p.bin_start + cell), e = __ldg(p.bin_start + cell + 1);
      for (int k = s; k < e; ++k) {
        const int4 en = __ldg(p.entries + k);
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
  size_t ghat, cnt, bin_cnt, bin_start, scan_state, counter, entries, total, zero_bytes;
  int nchunks;
  int64_t M;
};

static BwdWs bwd_ws_layout(int64_t N, int B, int V, int C, int H, int W) {
  BwdWs w;
  w.M = (int64_t)V * B * H * W;
  w.nchunks = (int)((w.M + kScanChunk - 1) / kScanChunk);
  size_t o = 0;
  const size_t n1 = (size_t)(N > 0 ? N : 1);
  w.ghat = o; o = align_up(o + sizeof(float) * n1 * (size_t)C, 256);
  w.cnt = o; o = align_up(o + sizeof(float) * n1, 256);
  w.bin_cnt = o; o = align_up(o + sizeof(int) * (size_t)w.M, 256);
  w.scan_state = o; o = align_up(o + sizeof(unsigned long long) * (size_t)(w.nchunks + 1), 256);
  w.counter = o; o = align_up(o + 256, 256);
  w.zero_bytes = o - w.bin_cnt;  // bin_cnt, scan_state and the ticket are cleared by ONE memset
  w.bin_start = o; o = align_up(o + sizeof(int) * (size_t)(w.M + 1), 256);
  w.entries = o; o = align_up(o + sizeof(int4) * n1 * (size_t)V, 256);
  w.total = o;
  return w;
}

typedef void (*gather_kernel_t)(const BwdParams);
static gather_kernel_t pick_gather_kernel(int C, int& NG) {
  NG = 1;
  if (C % 4 != 0) return nullptr;
  const int q = C / 4;
#define D3M_BWD_CASE(g, r)              \
  if (q == (g) * (r)) {                 \
    NG = 32 / (g);                      \
    return bp_bwd_gather_kernel<g, r>;  \
  }
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

template <int KIND>
static int launch_bwd(const BwdParams& p, size_t zero_bytes, float* cnt_ws, cudaStream_t stream) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  D3M_CUDA_CHECK(cudaMemsetAsync(p.bin_cnt, 0, zero_bytes, stream));
  const unsigned vox_ctas = (unsigned)((p.N + kSampleThreads - 1) / kSampleThreads);
  if (cnt_ws) {  // no forward count handed over: recompute it
    LaunchScope ls("bp_bwd_count", stream);
    bp_bwd_count_kernel<KIND><<<vox_ctas, kSampleThreads, 0, stream>>>(p, cnt_ws);
  }
  D3M_CUDA_CHECK(cudaGetLastError());
  {
    const int64_t ghat_blocks = (p.N * p.C + kSampleThreads * kGhatPerThread - 1) / (kSampleThreads * kGhatPerThread);
    const unsigned ghat_rows = (unsigned)((ghat_blocks + vox_ctas - 1) / vox_ctas);
    LaunchScope ls("bp_bwd_hist_ghat", stream);
    bp_bwd_sample_kernel<KIND, false><<<dim3(vox_ctas, p.V + ghat_rows), kSampleThreads, 0, stream>>>(p);
  }
  D3M_CUDA_CHECK(cudaGetLastError());
  {
    LaunchScope ls("bp_bwd_scan", stream);
    bp_scan_kernel<<<p.nchunks, kScanThreads, 0, stream>>>(p);
  }
  D3M_CUDA_CHECK(cudaGetLastError());
  {
    LaunchScope ls("bp_bwd_fill", stream);
    bp_bwd_sample_kernel<KIND, true><<<dim3(vox_ctas, p.V), kSampleThreads, 0, stream>>>(p);
  }
  D3M_CUDA_CHECK(cudaGetLastError());
  {
    // cells per warp task: small enough that even the coarsest level spreads over every SM
    int bins = (int)(p.M / ((int64_t)sms * 32));
    bins = bins < 8 ? 8 : (bins > 256 ? 256 : bins);
    const int64_t ntasks = (p.M + bins - 1) / bins;
    int64_t ctas = (ntasks + 7) / 8;
    if (ctas > (int64_t)sms * 64) ctas = (int64_t)sms * 64;
    if (ctas < 1) ctas = 1;
    LaunchScope ls("bp_bwd_order", stream);
    bp_bwd_order_kernel<<<(unsigned)ctas, 256, 0, stream>>>(p, bins);
    D3M_CUDA_CHECK(cudaGetLastError());
  }
  {
    int NG;
    gather_kernel_t k = pick_gather_kernel(p.C, NG);
    if (!k) {
      D3M_REQUIRE(p.C <= 256, D3M_ERR_ARG, "back_project backward: C=%d unsupported (C%%4!=0 needs C<=256)", p.C);
      k = bp_bwd_gather_generic_kernel;
      NG = 1;
    }
    const int64_t nsteps = (p.M + NG - 1) / NG;
    int64_t ctas = (nsteps + kGatherWarps - 1) / kGatherWarps;
    if (ctas > (int64_t)sms * 16) ctas = (int64_t)sms * 16;
    if (ctas < 1) ctas = 1;
    LaunchScope ls("bp_bwd_gather", stream);
    k<<<(unsigned)ctas, kGatherWarps * 32, 0, stream>>>(p);
    D3M_CUDA_CHECK(cudaGetLastError());
  }
  return D3M_OK;
}

}  // namespace d3m

using namespace d3m;

extern "C" size_t d3m_back_project_bwd_workspace(int64_t N, int B, int V, int C, int H, int W) {
  if (N < 0 || B < 1 || V < 1 || C < 1 || H < 1 || W < 1) return 0;
  return bwd_ws_layout(N, B, V, C, H, W).total;
}

extern "C" int d3m_back_project_bwd(const void* coords, int coords_kind, int64_t N, const float* origin, int B,
                                    float voxel_size, int V, int C, int H, int W, const float* KRcam,
                                    const float* grad_out, const float* count, float* grad_feats_nhwc, int grad_nchw,
                                    void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  D3M_REQUIRE(d3m_device_count() > 0, D3M_ERR_NO_DEVICE,
              "back_project backward: no CUDA device (there is no CPU fallback)");
  D3M_REQUIRE(N >= 0 && B >= 1 && V >= 1 && C >= 1 && H >= 2 && W >= 2, D3M_ERR_ARG,
              "back_project backward: bad sizes N=%lld B=%d V=%d C=%d H=%d W=%d", (long long)N, B, V, C, H, W);
  D3M_REQUIRE(coords_kind >= 0 && coords_kind <= 2, D3M_ERR_ARG, "back_project backward: coords_kind=%d", coords_kind);
  D3M_REQUIRE((int64_t)V * B * H * W < (1ll << 30), D3M_ERR_ARG, "back_project backward: V*B*H*W must be < 2^30");
  D3M_REQUIRE(V <= 65535 - 4096, D3M_ERR_ARG, "back_project backward: too many views");
  D3M_REQUIRE(N * (int64_t)V < (1ll << 31), D3M_ERR_ARG, "back_project backward: N*V must be < 2^31 samples");
  D3M_REQUIRE(grad_feats_nhwc && workspace, D3M_ERR_ARG, "back_project backward: NULL pointer");
  const BwdWs w = bwd_ws_layout(N, B, V, C, H, W);
  if (N == 0) {
    D3M_CUDA_CHECK(cudaMemsetAsync(grad_feats_nhwc, 0, sizeof(float) * (size_t)w.M * C, stream));
    return D3M_OK;
  }
  D3M_REQUIRE(coords && origin && KRcam && grad_out, D3M_ERR_ARG, "back_project backward: NULL pointer");
  D3M_REQUIRE(aligned16(coords) && aligned16(KRcam) && aligned16(grad_feats_nhwc) && aligned16(workspace),
              D3M_ERR_ALIGN, "back_project backward: coords/KRcam/grad_feats/workspace must be 16-byte aligned");
  D3M_REQUIRE(workspace_bytes >= w.total, D3M_ERR_WORKSPACE, "back_project backward: workspace %zu < %zu",
              workspace_bytes, w.total);
  unsigned char* ws = static_cast<unsigned char*>(workspace);
  BwdParams p;
  p.coords = coords; p.N = N; p.origin = origin; p.B = B; p.vs = voxel_size;
  p.V = V; p.C = C; p.H = H; p.W = W; p.KR = KRcam; p.grad_out = grad_out;
  p.ghat = reinterpret_cast<float*>(ws + w.ghat);
  p.bin_cnt = reinterpret_cast<int*>(ws + w.bin_cnt);
  p.bin_start = reinterpret_cast<int*>(ws + w.bin_start);
  p.scan_state = reinterpret_cast<unsigned long long*>(ws + w.scan_state);
  p.counter = reinterpret_cast<unsigned int*>(ws + w.counter);
  p.entries = reinterpret_cast<int4*>(ws + w.entries);
  p.grad_feats = grad_feats_nhwc;
  p.M = w.M;
  p.nchunks = w.nchunks;
  p.grad_nchw = grad_nchw ? 1 : 0;
  float* cnt_ws = count ? nullptr : reinterpret_cast<float*>(ws + w.cnt);
  p.count = count ? count : cnt_ws;
  if (coords_kind == D3M_COORDS_F32) return launch_bwd<D3M_COORDS_F32>(p, w.zero_bytes, cnt_ws, stream);
  if (coords_kind == D3M_COORDS_I64) return launch_bwd<D3M_COORDS_I64>(p, w.zero_bytes, cnt_ws, stream);
  return launch_bwd<D3M_COORDS_I32>(p, w.zero_bytes, cnt_ws, stream);
}
