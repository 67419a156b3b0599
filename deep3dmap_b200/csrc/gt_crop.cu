// SURVEY §8 row f1 (second half) -- the ground-truth side of the dataloader transform
// `SeqRandomTransformSpace.transform` (datasets/pipelines/transforms_seq.py:343-398): after the nine views of a
// fragment were integrated into a small TSDFVolumeTorch volume (csrc/tsdf.cu), the reference
//   * thresholds that volume into the occupancy GT                                  (:365-366)
//   * re-samples the full-scene TSDF of level l on the transformed fragment grid with TWO 3-D grid_sample calls
//     (nearest and trilinear, align_corners=False, zero padding), takes the trilinear value where |nearest| < 1 and
//     marks every voxel whose normalised coordinate leaves [-1,1) as empty (= 1)     (:368-396)
// materialising the (3, 96^3) coordinate tensor four times on the way.  Here both are one streaming kernel each: the
// coordinate chain is evaluated per voxel in registers with the reference's fp32 rounding sequence (explicit _rn
// intrinsics, no contraction), the nine taps come from the L2-resident scene volume.
#include "d3m_common.cuh"

namespace d3m {

struct CropParams {
  const float* full;   // (X, Y, Z) scene TSDF of this level, C order
  int X, Y, Z;
  int nx, ny, nz;      // output dims = voxel_dim / step
  int step;            // 2^l
  float voxel_size;    // finest voxel size (the reference's self.voxel_size)
  float inv_step;      // 1 / 2^l (exact)
  float opx, opy, opz; // vol_origin_partial
  float oox, ooy, ooz; // old_origin
  float t[12];         // transform[:3, :], row-major
  float* out;          // (nx, ny, nz)
};

// aten grid_sampler_unnormalize, align_corners = False:  ((coord + 1) * size - 1) / 2
__device__ __forceinline__ float unnormalize(float g, int size) {
  return __fmul_rn(__fsub_rn(__fmul_rn(__fadd_rn(g, 1.0f), (float)size), 1.0f), 0.5f);
}

__device__ __forceinline__ float tap(const CropParams& p, int ix, int iy, int iz) {
  // ix indexes Z (W of the 5-D input), iy indexes Y (H), iz indexes X (D); zero padding
  if (ix < 0 || ix >= p.Z || iy < 0 || iy >= p.Y || iz < 0 || iz >= p.X) return 0.0f;
  return __ldg(p.full + ((int64_t)iz * p.Y + iy) * p.Z + ix);
}

__global__ void __launch_bounds__(256) gt_recrop_kernel(const CropParams p) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t total = (int64_t)p.nx * p.ny * p.nz;
  if (n >= total) return;
  const int k = (int)(n % p.nz);
  const int64_t r = n / p.nz;
  const int j = (int)(r % p.ny);
  const int i = (int)(r / p.ny);
  // :346-347  world = coords.float() * voxel_size + vol_origin_partial   (coords of the finest grid, strided by step)
  const float wx = __fadd_rn(__fmul_rn((float)(i * p.step), p.voxel_size), p.opx);
  const float wy = __fadd_rn(__fmul_rn((float)(j * p.step), p.voxel_size), p.opy);
  const float wz = __fadd_rn(__fmul_rn((float)(k * p.step), p.voxel_size), p.opz);
  // :349  transform[:3, :] @ [world; 1]   (K = 4 dot product as sgemm's k = 0..3 FMA chain)
  float c[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float* t = p.t + 4 * a;
    const float v = __fmaf_rn(t[3], 1.0f, __fmaf_rn(t[2], wz, __fmaf_rn(t[1], wy, __fmul_rn(t[0], wx))));
    const float oo = a == 0 ? p.oox : a == 1 ? p.ooy : p.ooz;
    // :350 (world - old_origin) / voxel_size ; :370 / 2**l ; :377 2 * c / (old_dim - 1) - 1
    const float cs = __fmul_rn(__fdiv_rn(__fsub_rn(v, oo), p.voxel_size), p.inv_step);
    const int dim = a == 0 ? p.X : a == 1 ? p.Y : p.Z;
    c[a] = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, cs), (float)(dim - 1)), 1.0f);
  }
  // :378 grid = coords[[2,1,0]] -> grid x walks the Z axis, grid z walks the X axis
  const float gx = c[2], gy = c[1], gz = c[0];
  float res;
  if (fabsf(gx) >= 1.0f || fabsf(gy) >= 1.0f || fabsf(gz) >= 1.0f) {
    res = 1.0f;  // :395-396 beyond the scene volume: empty.  (NaN coordinates compare false, as in torch)
  } else {
    const float ix = unnormalize(gx, p.Z), iy = unnormalize(gy, p.Y), iz = unnormalize(gz, p.X);
    // mode='nearest': std::nearbyint = round half to even
    const float near = tap(p, __float2int_rn(ix), __float2int_rn(iy), __float2int_rn(iz));
    res = near;
    if (fabsf(near) < 1.0f) {
      // mode='bilinear' (trilinear), aten's corner order and weight expressions
      const float fx0 = floorf(ix), fy0 = floorf(iy), fz0 = floorf(iz);
      const int x0 = (int)fx0, y0 = (int)fy0, z0 = (int)fz0;
      const float fx1 = (float)(x0 + 1), fy1 = (float)(y0 + 1), fz1 = (float)(z0 + 1);
      const float ax1 = __fsub_rn(fx1, ix), ax0 = __fsub_rn(ix, fx0);
      const float ay1 = __fsub_rn(fy1, iy), ay0 = __fsub_rn(iy, fy0);
      const float az1 = __fsub_rn(fz1, iz), az0 = __fsub_rn(iz, fz0);
      const float w_tnw = __fmul_rn(__fmul_rn(ax1, ay1), az1);
      const float w_tne = __fmul_rn(__fmul_rn(ax0, ay1), az1);
      const float w_tsw = __fmul_rn(__fmul_rn(ax1, ay0), az1);
      const float w_tse = __fmul_rn(__fmul_rn(ax0, ay0), az1);
      const float w_bnw = __fmul_rn(__fmul_rn(ax1, ay1), az0);
      const float w_bne = __fmul_rn(__fmul_rn(ax0, ay1), az0);
      const float w_bsw = __fmul_rn(__fmul_rn(ax1, ay0), az0);
      const float w_bse = __fmul_rn(__fmul_rn(ax0, ay0), az0);
      float acc = 0.0f;
      acc = __fadd_rn(acc, __fmul_rn(tap(p, x0, y0, z0), w_tnw));
      acc = __fadd_rn(acc, __fmul_rn(tap(p, x0 + 1, y0, z0), w_tne));
      acc = __fadd_rn(acc, __fmul_rn(tap(p, x0, y0 + 1, z0), w_tsw));
      acc = __fadd_rn(acc, __fmul_rn(tap(p, x0 + 1, y0 + 1, z0), w_tse));
      acc = __fadd_rn(acc, __fmul_rn(tap(p, x0, y0, z0 + 1), w_bnw));
      acc = __fadd_rn(acc, __fmul_rn(tap(p, x0 + 1, y0, z0 + 1), w_bne));
      acc = __fadd_rn(acc, __fmul_rn(tap(p, x0, y0 + 1, z0 + 1), w_bsw));
      acc = __fadd_rn(acc, __fmul_rn(tap(p, x0 + 1, y0 + 1, z0 + 1), w_bse));
      res = acc;
    }
  }
  __stcs(p.out + n, res);
}

// :365-366  occ = (tsdf < hi) & (tsdf > lo) & (weight > min_weight); four voxels per thread
__global__ void __launch_bounds__(256) tsdf_occupancy_kernel(const float* __restrict__ tsdf, const float* __restrict__ weight,
                                                             int64_t n, float lo, float hi, float min_weight,
                                                             uint8_t* __restrict__ occ, bool vec) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  if (vec && i + 4 <= n) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(tsdf + i));
    const float4 w = __ldg(reinterpret_cast<const float4*>(weight + i));
    uchar4 o;
    o.x = (t.x < hi && t.x > lo && w.x > min_weight) ? 1 : 0;
    o.y = (t.y < hi && t.y > lo && w.y > min_weight) ? 1 : 0;
    o.z = (t.z < hi && t.z > lo && w.z > min_weight) ? 1 : 0;
    o.w = (t.w < hi && t.w > lo && w.w > min_weight) ? 1 : 0;
    *reinterpret_cast<uchar4*>(occ + i) = o;
  } else {
    for (int64_t q = i; q < n && q < i + 4; ++q) {
      const float t = __ldg(tsdf + q);
      occ[q] = (t < hi && t > lo && __ldg(weight + q) > min_weight) ? 1 : 0;
    }
  }
}

}  // namespace d3m

using namespace d3m;

extern "C" int d3m_gt_recrop(const float* tsdf_full, int X, int Y, int Z, int nx, int ny, int nz, int step,
                             float voxel_size, const float* vol_origin_partial3_host, const float* transform12_host,
                             const float* old_origin3_host, float* out, void* stream_) {
  D3M_REQUIRE(d3m_device_count() > 0, D3M_ERR_NO_DEVICE, "d3m_gt_recrop: no CUDA device (there is no CPU fallback)");
  D3M_REQUIRE(X >= 1 && Y >= 1 && Z >= 1 && nx >= 0 && ny >= 0 && nz >= 0 && step >= 1 && (step & (step - 1)) == 0 &&
                  voxel_size > 0.0f,
              D3M_ERR_ARG, "d3m_gt_recrop: bad arguments (step must be a power of two)");
  D3M_REQUIRE(vol_origin_partial3_host && transform12_host && old_origin3_host, D3M_ERR_ARG,
              "d3m_gt_recrop: NULL host parameter");
  const int64_t total = (int64_t)nx * ny * nz;
  if (total == 0) return D3M_OK;
  D3M_REQUIRE(tsdf_full && out, D3M_ERR_ARG, "d3m_gt_recrop: NULL pointer");
  CropParams p;
  p.full = tsdf_full;
  p.X = X; p.Y = Y; p.Z = Z;
  p.nx = nx; p.ny = ny; p.nz = nz;
  p.step = step;
  p.voxel_size = voxel_size;
  p.inv_step = 1.0f / (float)step;
  p.opx = vol_origin_partial3_host[0]; p.opy = vol_origin_partial3_host[1]; p.opz = vol_origin_partial3_host[2];
  p.oox = old_origin3_host[0]; p.ooy = old_origin3_host[1]; p.ooz = old_origin3_host[2];
  for (int q = 0; q < 12; ++q) p.t[q] = transform12_host[q];
  p.out = out;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LaunchScope ls("gt_recrop", stream);
  gt_recrop_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(p);
  D3M_CUDA_CHECK(cudaGetLastError());
  return D3M_OK;
}

extern "C" int d3m_tsdf_occupancy(const float* tsdf, const float* weight, int64_t n, float lo, float hi,
                                  float min_weight, uint8_t* occ, void* stream_) {
  D3M_REQUIRE(d3m_device_count() > 0, D3M_ERR_NO_DEVICE, "d3m_tsdf_occupancy: no CUDA device (there is no CPU fallback)");
  D3M_REQUIRE(n >= 0, D3M_ERR_ARG, "d3m_tsdf_occupancy: bad arguments");
  if (n == 0) return D3M_OK;
  D3M_REQUIRE(tsdf && weight && occ, D3M_ERR_ARG, "d3m_tsdf_occupancy: NULL pointer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool vec = ((reinterpret_cast<uintptr_t>(tsdf) | reinterpret_cast<uintptr_t>(weight)) & 15u) == 0 &&
                   (reinterpret_cast<uintptr_t>(occ) & 3u) == 0;
  LaunchScope ls("tsdf_occupancy", stream);
  tsdf_occupancy_kernel<<<(unsigned)((n + 1023) / 1024), 256, 0, stream>>>(tsdf, weight, n, lo, hi, min_weight, occ, vec);
  D3M_CUDA_CHECK(cudaGetLastError());
  return D3M_OK;
}
