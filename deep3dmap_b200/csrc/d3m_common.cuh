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
bool pdl_enabled();  // core.cu: env D3M_PDL != "0"

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
