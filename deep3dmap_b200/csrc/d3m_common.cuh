// Shared device/host helpers of libd3m (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/d3m.h"

namespace d3m {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define D3M_CUDA_CHECK(expr)                                   \
  do {                                                         \
    cudaError_t _e = (expr);                                   \
    if (_e != cudaSuccess) return ::d3m::cuda_fail(_e, #expr); \
  } while (0)

#define D3M_REQUIRE(cond, code, ...)   \
  do {                                 \
    if (!(cond)) {                     \
      ::d3m::set_error(__VA_ARGS__);   \
      return (code);                   \
    }                                  \
  } while (0)

// RAII marker around one kernel launch: counts it and, while profiling is on, brackets it with CUDA events
// recorded on the launching stream (bench.py's live per-kernel timing).
struct LaunchScope {
  LaunchScope(const char* name, cudaStream_t stream);
  ~LaunchScope();
  const char* name_;
  cudaStream_t stream_;
  cudaEvent_t start_;
};

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Makes `device` current for the lifetime of the guard and restores the caller's device afterwards: handle-owning entry
// points (d3m_tsdf_*) must not leave the calling thread on another GPU (torch allocations would silently follow).
struct DeviceGuard {
  explicit DeviceGuard(int device) : prev_(-1), switched_(false), err_(cudaSuccess) {
    err_ = cudaGetDevice(&prev_);
    if (err_ == cudaSuccess && prev_ != device) {
      err_ = cudaSetDevice(device);
      switched_ = (err_ == cudaSuccess);
    }
  }
  ~DeviceGuard() {
    if (switched_) cudaSetDevice(prev_);
  }
  cudaError_t error() const { return err_; }
  int prev_;
  bool switched_;
  cudaError_t err_;
};

// ------------------------------------------------------------------------------------------------
// Programmatic dependent launch (griddepcontrol, sm_90+).  The fragment-sized step is a chain of 5-45 us kernels, so
// the launch ramp of kernel K+1 (CTA dispatch, parameter and instruction fetch) is a visible share of the step.  Every
// kernel of the chain starts with pdl_enter(): wait until the kernel before it has completed and its writes are
// visible, THEN allow the next kernel of the stream to be dispatched.  A successor launched through launch_k() therefore
// becomes resident while this kernel is still running (its CTAs block in griddepcontrol.wait and fill the SMs as ours
// drain) instead of after our last CTA has retired.  Look-ahead is one kernel deep by construction, and a successor is
// only dispatched once every CTA of its predecessor has started, so waiting CTAs can never starve running ones.
// Without the launch attribute (plain <<< >>>, D3M_PDL=0, or a memset / torch op in between) both instructions are no-ops
// and the stream serialises as usual.  Works in eager streams and inside CUDA-graph capture (programmatic edges).
// ------------------------------------------------------------------------------------------------
bool pdl_enabled();  // core.cu: D3M_PDL = 0 / 1, or (default) the decision of the enclosing PdlScope

// Per-call decision of the default mode: chained launches only for calls small enough to be launch-latency bound
// (voxel-view samples <= kPdlMaxWorkItems) and only outside CUDA-graph capture (see core.cu for the measurements).
constexpr long long kPdlMaxWorkItems = 4ll << 20;
struct PdlScope {
  PdlScope(cudaStream_t stream, long long work_items);
  ~PdlScope();
  bool prev_;
};

__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                   Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// host-side caches of things every call used to ask the driver (an eager fragment step is host-bound):
int current_device_sms();                                  // SM count of the current device
cudaError_t ensure_dynamic_smem(const void* kernel, size_t bytes);  // cudaFuncSetAttribute only when the need grows

// zero-fill on the stream: a PDL-chained kernel (core.cu) when it can be, cudaMemsetAsync otherwise
int zero_async(void* p, size_t bytes, cudaStream_t stream);

// ------------------------------------------------------------------------------------------------
// Backward binning.  A bin = (bilinear cell (v,b,y0,x0), voxel bucket n & (nb-1)): bins are ordered cell-major,
// bucket-minor, so a cell's entries stay contiguous; inside a bin the entries are ranked by voxel index.  The resulting
// accumulation order per cell -- (low bits of the voxel index, voxel index) -- is a pure function of the inputs, which
// is all determinism needs.  More buckets (nb = 2^nb_log2) shorten the bins of heavily populated cells (a distant wall
// of a large scene lands thousands of voxels in one cell and ranking is quadratic in the bin length); the LOW bits
// are used because the voxels that share a cell are spatially clustered, i.e. share their high index bits.  The price
// is nb x larger scan arrays, so nb grows with the number of potential samples per cell.
// Pure function of the call's shapes: forward (histogram) and backward (scan / fill / order / gather) agree on it.
// ------------------------------------------------------------------------------------------------
struct BinCfg {
  int nb_log2;  // buckets per cell = 1 << nb_log2; bucket = voxel index & ((1 << nb_log2) - 1)
};
static inline BinCfg bin_config(int64_t N, int V, int64_t M) {
  BinCfg c;
  c.nb_log2 = 0;
  const double per_cell = M > 0 ? (double)N * V / (double)M : 0.0;
  while (c.nb_log2 < 6 && per_cell > 32.0 * (double)(1 << c.nb_log2) && (M << (c.nb_log2 + 1)) < (1ll << 30)) ++c.nb_log2;
  return c;
}

// ------------------------------------------------------------------------------------------------
// Binning state: ONE int32 buffer that travels from the forward call to the backward call (the `cell_hist` argument of
// the C ABI) or lives in the backward workspace when forward did not produce it.  Offsets in int32 elements, every
// section 16-byte aligned:
//   [cnt      Mb      ]  samples per bin, integer RED in bp_fwd (or bp_bwd_hist)            zeroed by the prep kernel
//   [cursor   Mb      ]  claim counters of the fill pass; gather re-zeroes every cell it owns  (self-cleaning)
//   [scan_st  2*nchunks]  decoupled-look-back words; the last scan CTA re-zeroes them          (self-cleaning)
//   [counters 16      ]  0: scan ticket, 1: scan done, 2: forward stats ticket                (self-cleaning)
//   ---------------------  zero_elems: the prefix the prep kernel clears ONCE per forward call
//   [start    Mb+1    ]  exclusive scan of cnt (written by the scan CTAs of bp_fwd_finish)
// Because cursor / scan state / counters clean up after themselves, backward may run any number of times on the state of
// one forward call without another clear, and no launch of the fragment step exists only to zero memory.
// ------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanChunk = kScanThreads * kScanItems;
enum { kCtrScanTicket = 0, kCtrScanDone = 1, kCtrStatsTicket = 2, kCtrCount = 16 };

struct BinLayout {
  int64_t M, Mb;
  int nb_log2, nchunks;
  size_t cnt, cursor, scan_state, counters, zero_elems, start, total;  // int32 element offsets
};
static inline size_t align4(size_t x) { return (x + 3) & ~(size_t)3; }
static inline BinLayout bin_layout(int64_t N, int B, int V, int H, int W) {
  BinLayout l;
  l.M = (int64_t)V * B * H * W;
  l.nb_log2 = bin_config(N, V, l.M).nb_log2;
  l.Mb = l.M << l.nb_log2;
  l.nchunks = (int)((l.Mb + kScanChunk - 1) / kScanChunk);
  size_t o = 0;
  l.cnt = o; o = align4(o + (size_t)l.Mb);
  l.cursor = o; o = align4(o + (size_t)l.Mb);
  l.scan_state = o; o = align4(o + 2 * (size_t)l.nchunks);
  l.counters = o; o = align4(o + kCtrCount);
  l.zero_elems = o;
  l.start = o; o = align4(o + (size_t)l.Mb + 1);
  l.total = o;
  return l;
}

// Device view of the binning state.
struct BinState {
  int* cnt;
  int* cursor;
  int* start;                      // Mb + 1
  unsigned long long* scan_state;  // nchunks
  unsigned int* counters;
  int64_t Mb;
  int nb_log2, nchunks;
};
static inline BinState bin_state(int* base, const BinLayout& l) {
  BinState s;
  s.cnt = base + l.cnt; s.cursor = base + l.cursor; s.start = base + l.start;
  s.scan_state = reinterpret_cast<unsigned long long*>(base + l.scan_state);
  s.counters = reinterpret_cast<unsigned int*>(base + l.counters);
  s.Mb = l.Mb; s.nb_log2 = l.nb_log2; s.nchunks = l.nchunks;
  return s;
}

// ---- exclusive scan of the bin histogram: ONE pass, chained look-back ---------------------------------------------
// Executed by a whole CTA of kScanThreads threads for ONE chunk.  CTAs take chunk ids from a ticket (so a chunk's
// predecessors are always already scheduled), publish their aggregate, walk back over predecessors until they meet an
// inclusive prefix, then publish their own.  state word = (flag << 32) | value, flag 0 = not ready, 1 = aggregate,
// 2 = inclusive prefix.  Integer sums: the result does not depend on the order in which CTAs arrive.  The last CTA to
// finish clears the look-back words and both counters again (see BinLayout).
__device__ __forceinline__ void scan_bins_cta(const BinState& st) {
  __shared__ int red[kScanThreads / 32];
  __shared__ int s_cid, s_prefix;
  __shared__ bool s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) s_cid = (int)atomicAdd(&st.counters[kCtrScanTicket], 1u);
  __syncthreads();
  const int cid = s_cid;
  const int64_t base = (int64_t)cid * kScanChunk + (int64_t)tid * kScanItems;
  int v[kScanItems];
  if (base + kScanItems <= st.Mb) {
    const int4* q = reinterpret_cast<const int4*>(st.cnt + base);
#pragma unroll
    for (int i = 0; i < kScanItems / 4; ++i) {
      const int4 t = __ldcg(q + i);
      v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) v[i] = (base + i < st.Mb) ? __ldcg(st.cnt + base + i) : 0;
  }
  int s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) s += v[i];
  int inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
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
  if (warp == 0) {
    // decoupled look-back, one warp wide: lane l inspects chunk cid-1-l; the nearest predecessor that already holds an
    // inclusive prefix (state 2) ends the walk, the aggregates (state 1) in front of it are summed with one shuffle tree.
    volatile unsigned long long* sw = st.scan_state;
    int prefix = 0;
    if (cid == 0) {
      if (lane == 0) sw[0] = (2ull << 32) | (unsigned)total;
    } else {
      if (lane == 0) {
        sw[cid] = (1ull << 32) | (unsigned)total;
        __threadfence();
      }
      for (int j0 = cid - 1;; j0 -= 32) {
        const int j = j0 - lane;
        unsigned long long w = 2ull << 32;  // "chunk -1": inclusive prefix 0
        if (j >= 0) {
          do { w = sw[j]; } while ((w >> 32) == 0ull);
        }
        const unsigned done = __ballot_sync(0xffffffffu, (w >> 32) == 2ull);
        const int first = __ffs(done) - 1;  // -1: no inclusive prefix in this window
        int use = (done == 0u || lane <= first) ? (int)(unsigned)(w & 0xffffffffull) : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) use += __shfl_xor_sync(0xffffffffu, use, o);
        prefix += use;
        if (done != 0u) break;
      }
      if (lane == 0) sw[cid] = (2ull << 32) | (unsigned)(prefix + total);
    }
    if (lane == 0) {
      s_prefix = prefix;
      if (cid == st.nchunks - 1) st.start[st.Mb] = prefix + total;
    }
  }
  __syncthreads();
  int off = s_prefix + woff + inc - s;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    if (base + i < st.Mb) st.start[base + i] = off;
    off += v[i];
  }
  // self-cleaning: once every chunk has finished its look-back nobody reads the state words any more
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    s_last = atomicAdd(&st.counters[kCtrScanDone], 1u) == (unsigned)(st.nchunks - 1);
  }
  __syncthreads();
  if (s_last) {
    for (int i = tid; i < st.nchunks; i += kScanThreads) st.scan_state[i] = 0ull;
    if (tid == 0) { st.counters[kCtrScanTicket] = 0u; st.counters[kCtrScanDone] = 0u; }
  }
}

__host__ __device__ static inline int align_up_dev(int x) { return (x + 15) & ~15; }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ------------------------------------------------------------------------------------------------
// coords row -> (fragment index or -1, xyz as float).  back_project.py:29-30: a row belongs to
// fragment b iff coords[:,0] == b.
// ------------------------------------------------------------------------------------------------
template <int KIND>
__device__ __forceinline__ int load_coord(const void* __restrict__ coords, int64_t n, int B, float& x, float& y,
                                          float& z) {
  int b = -1;
  if (KIND == D3M_COORDS_F32) {
    const float4 c = __ldg(reinterpret_cast<const float4*>(coords) + n);
    if (c.x >= 0.0f && c.x < (float)B && c.x == floorf(c.x)) b = (int)c.x;
    x = c.y; y = c.z; z = c.w;
  } else if (KIND == D3M_COORDS_I64) {
    const longlong2 a = __ldg(reinterpret_cast<const longlong2*>(coords) + 2 * n);
    const longlong2 c = __ldg(reinterpret_cast<const longlong2*>(coords) + 2 * n + 1);
    if (a.x >= 0 && a.x < (long long)B) b = (int)a.x;
    x = (float)a.y; y = (float)c.x; z = (float)c.y;
  } else {
    const int4 c = __ldg(reinterpret_cast<const int4*>(coords) + n);
    if (c.x >= 0 && c.x < B) b = c.x;
    x = (float)c.y; y = (float)c.z; z = (float)c.w;
  }
  return b;
}

// ------------------------------------------------------------------------------------------------
// One voxel-view projection with the exact fp32 rounding sequence of the reference
// (back_project.py:37-51 + aten grid_sampler un-normalisation, align_corners=True):
//   grid = coords*vs + origin (mul, add) ; p_r = fma chain over k=0..3 (== torch bmm, K=4)
//   u = p0/p2 ; nx = (2u)/(W-1) - 1 ; valid = |nx|<=1 & |ny|<=1 & p2>0
//   ix = ((nx+1)/2)*(W-1) ; x0 = floor(ix) ; fx = ix-x0 (exact)
// Explicit _rn intrinsics keep nvcc from contracting anything.
// ------------------------------------------------------------------------------------------------
struct Sample {
  float z, fx, fy;
  int x0, y0;
  bool valid;
};

__device__ __forceinline__ void voxel_world(float cx, float cy, float cz, float vs, float ox, float oy, float oz,
                                            float& gx, float& gy, float& gz) {
  gx = __fadd_rn(__fmul_rn(cx, vs), ox);
  gy = __fadd_rn(__fmul_rn(cy, vs), oy);
  gz = __fadd_rn(__fmul_rn(cz, vs), oz);
}

// P: first three rows of the 4x4 matrix, row-major (12 floats used of 16)
__device__ __forceinline__ Sample project(float gx, float gy, float gz, const float4 r0, const float4 r1,
                                          const float4 r2, float wm1, float hm1) {
  Sample s;
  const float p0 = __fadd_rn(__fmaf_rn(r0.z, gz, __fmaf_rn(r0.y, gy, __fmul_rn(r0.x, gx))), r0.w);
  const float p1 = __fadd_rn(__fmaf_rn(r1.z, gz, __fmaf_rn(r1.y, gy, __fmul_rn(r1.x, gx))), r1.w);
  const float p2 = __fadd_rn(__fmaf_rn(r2.z, gz, __fmaf_rn(r2.y, gy, __fmul_rn(r2.x, gx))), r2.w);
  // Early outs before the four IEEE divisions (most samples of a large scene are far outside the image).  Both are
  // exact with respect to the test below: `valid` needs p2 > 0; and for p2 > 0 a pixel coordinate more than 0.1 % of the
  // image size outside [0, W-1] x [0, H-1] cannot round back inside (the rounding error of the chain is ~1e-6 relative).
  // Borderline samples fall through to the reference arithmetic, so count / masks stay bit-exact.
  s.z = p2; s.fx = 0.0f; s.fy = 0.0f; s.x0 = 0; s.y0 = 0; s.valid = false;
  if (!(p2 > 0.0f)) return s;
  {
    const float mx = __fmul_rn(1e-3f, wm1), my = __fmul_rn(1e-3f, hm1);
    if (p0 < -mx * p2 || p0 > (wm1 + mx) * p2 || p1 < -my * p2 || p1 > (hm1 + my) * p2) return s;
  }
  const float u = __fdiv_rn(p0, p2);
  const float v = __fdiv_rn(p1, p2);
  const float nx = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, u), wm1), 1.0f);
  const float ny = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, v), hm1), 1.0f);
  s.valid = (fabsf(nx) <= 1.0f) && (fabsf(ny) <= 1.0f) && (p2 > 0.0f);
  const float ix = __fmul_rn(__fmul_rn(__fadd_rn(nx, 1.0f), 0.5f), wm1);
  const float iy = __fmul_rn(__fmul_rn(__fadd_rn(ny, 1.0f), 0.5f), hm1);
  const float x0f = floorf(ix), y0f = floorf(iy);
  s.x0 = (int)x0f;
  s.y0 = (int)y0f;
  s.fx = __fsub_rn(ix, x0f);
  s.fy = __fsub_rn(iy, y0f);
  return s;
}

__device__ __forceinline__ void load_krcam(const float* __restrict__ KR, int v, int B, int b, float4& r0, float4& r1,
                                           float4& r2) {
  const float4* p = reinterpret_cast<const float4*>(KR) + ((int64_t)v * B + b) * 4;
  r0 = __ldg(p);
  r1 = __ldg(p + 1);
  r2 = __ldg(p + 2);
}

}  // namespace d3m
