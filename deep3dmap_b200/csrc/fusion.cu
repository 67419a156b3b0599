// SURVEY §8 f3 -- sparse <-> dense movement of the GRU-fusion global volume and the direct-substitute TSDF fuse
// (models/modulars/gru_fusion.py:51-181, core/utils/neucon_utils.py:114-131).  The reference expresses these steps
// as torch.full + index_put, boolean masks, torch.nonzero over the dense fragment-bounding volume (FBV) and advanced
// indexing; here each is one streaming kernel over the (X,Y,Z[,c]) volume or the coordinate list.  Ordered outputs
// (nonzero) reuse the compaction of level_glue.cu, so row order equals the reference's.
#include "d3m_common.cuh"

namespace d3m {

static inline unsigned nblocks(int64_t n, int per_block) { return (unsigned)((n + per_block - 1) / per_block); }

__global__ void __launch_bounds__(256) fill_kernel(float* __restrict__ p, int64_t n, float v) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i4 = i * 4;
  if (i4 + 3 < n) {
    reinterpret_cast<float4*>(p)[i] = make_float4(v, v, v, v);
  } else {
    for (int64_t k = i4; k < n; ++k) p[k] = v;
  }
}

__device__ __forceinline__ int64_t lin3(const int64_t* __restrict__ locs, int64_t m, int X, int Y, int Z) {
  long long x = __ldg(locs + 3 * m), y = __ldg(locs + 3 * m + 1), z = __ldg(locs + 3 * m + 2);
  // torch advanced indexing wraps negative indices once
  if (x < 0) x += X;
  if (y < 0) y += Y;
  if (z < 0) z += Z;
  if (x < 0 || x >= X || y < 0 || y >= Y || z < 0 || z >= Z) return -1;
  return (x * Y + y) * (int64_t)Z + z;
}

// pass 1 of the scatter: the LAST row that addresses a voxel owns it (index_put on the CPU reference applies rows
// in order, so the last duplicate wins; torch-CUDA leaves the winner unspecified)
__global__ void __launch_bounds__(256) scatter_owner_kernel(const int64_t* __restrict__ locs, int64_t M, int X, int Y,
                                                            int Z, int* __restrict__ owner, int* __restrict__ bad) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const int64_t lin = lin3(locs, m, X, Y, Z);
  if (lin < 0) {
    atomicAdd(bad, 1);
    return;
  }
  atomicMax(owner + lin, (int)m);
}

// pass 2: thread per (row, channel)
__global__ void __launch_bounds__(256) scatter_rows_kernel(const int64_t* __restrict__ locs, int64_t M, int X, int Y,
                                                           int Z, const int* __restrict__ owner,
                                                           const float* __restrict__ values, int value_rows, int c,
                                                           float scalar, float* __restrict__ dense) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M * c) return;
  const int64_t m = t / c;
  const int j = (int)(t - m * c);
  const int64_t lin = lin3(locs, m, X, Y, Z);
  if (lin < 0) return;
  if (owner && __ldg(owner + lin) != (int)m) return;
  dense[lin * c + j] = value_rows ? __ldg(values + m * c + j) : scalar;
}

// FBV membership of the global map's voxels (gru_fusion.py:83-91):
//   shifted = global_coords - relative_origin ; valid = all(0 <= shifted < dim)
//   and, when the sparsity is NOT re-derived (FUSION.FULL False), also "the current fragment has that voxel".
__global__ void __launch_bounds__(256) fbv_mask_kernel(const int64_t* __restrict__ coords, int64_t M, long long ox,
                                                       long long oy, long long oz, int X, int Y, int Z,
                                                       const float* __restrict__ occupied, int64_t* __restrict__ shifted,
                                                       uint8_t* __restrict__ valid) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const long long x = __ldg(coords + 3 * m) - ox, y = __ldg(coords + 3 * m + 1) - oy, z = __ldg(coords + 3 * m + 2) - oz;
  bool v = x >= 0 && x < X && y >= 0 && y < Y && z >= 0 && z < Z;
  if (v && occupied) v = __ldg(occupied + (x * Y + y) * (int64_t)Z + z) != 0.0f;
  if (shifted) {
    shifted[3 * m] = x;
    shifted[3 * m + 1] = y;
    shifted[3 * m + 2] = z;
  }
  valid[m] = v ? 1 : 0;
}

// sparsity of the fused fragment (gru_fusion.py:100-106):
//   mode 0:  (a != 0).any(-1) | (b != 0).any(-1)          feature volumes (default 0)
//   mode 1:  (|a| < 1).any(-1) | (|b| < 1).any(-1)        tsdf volumes (default 1)
__device__ __forceinline__ bool union_pred(float v, int mode) { return mode ? fabsf(v) < 1.0f : v != 0.0f; }

// c == 1 (tsdf volumes) and the generic fallback: one thread per voxel
__global__ void __launch_bounds__(256) union_flags_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                          int64_t n_vox, int c, int mode, uint8_t* __restrict__ flags) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_vox) return;
  bool f = false;
  for (int j = 0; j < c; ++j) {
    f = f || union_pred(__ldg(a + i * c + j), mode);
    if (b) f = f || union_pred(__ldg(b + i * c + j), mode);
  }
  flags[i] = f ? 1 : 0;
}

// c > 1: a warp owns 32 consecutive voxels = one contiguous span of 32*c floats of each volume and reads it with
// fully coalesced VecT loads (W = c / floats-per-VecT iterations, lane l takes unit k*32+l); the predicate of every
// unit goes through a ballot into a W-word bitmap per warp, and lane v then ORs the W bits of voxel v.  The
// thread-per-voxel version strode through memory 4*c bytes apart and reached 20 % of the HBM peak.
constexpr int UF_MAX_W = 64;
template <typename VecT>
__global__ void __launch_bounds__(256) union_flags_warp_kernel(const VecT* __restrict__ a, const VecT* __restrict__ b,
                                                               int64_t n_vox, int W, int mode,
                                                               uint8_t* __restrict__ flags) {
  __shared__ unsigned bits[8][UF_MAX_W];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t v0 = ((int64_t)blockIdx.x * 8 + wid) * 32;
  if (v0 >= n_vox) return;
  const int nv = (int)min((int64_t)32, n_vox - v0);
  const int n_units = nv * W;
  const VecT* pa = a + v0 * W;
  const VecT* pb = b ? b + v0 * W : nullptr;
#pragma unroll 4
  for (int k = 0; k < W; ++k) {
    const int u = k * 32 + lane;
    bool f = false;
    if (u < n_units) {
      const VecT x = __ldcs(pa + u);
      const float* xf = reinterpret_cast<const float*>(&x);
#pragma unroll
      for (int q = 0; q < (int)(sizeof(VecT) / 4); ++q) f = f || union_pred(xf[q], mode);
      if (pb) {
        const VecT y = __ldcs(pb + u);
        const float* yf = reinterpret_cast<const float*>(&y);
#pragma unroll
        for (int q = 0; q < (int)(sizeof(VecT) / 4); ++q) f = f || union_pred(yf[q], mode);
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, f);
    if (lane == 0) bits[wid][k] = m;
  }
  __syncwarp();
  if (lane < nv) {
    // bits [lane*W, lane*W + W) of the concatenated bitmap
    const int lo = lane * W, hi = lo + W;
    unsigned any = 0;
    for (int w = lo >> 5; w <= (hi - 1) >> 5; ++w) {
      unsigned m = bits[wid][w];
      const int b0 = w << 5;
      if (lo > b0) m &= 0xffffffffu << (lo - b0);
      if (hi < b0 + 32) m &= (1u << (hi - b0)) - 1u;
      any |= m;
    }
    flags[v0 + lane] = any ? 1 : 0;
  }
}

// linear voxel index -> (x,y,z) [+ offset] rows, optionally with a leading batch column and a scale on xyz
// (torch.nonzero on the dense volume, gru_fusion.py:104/106/145, and ":288  cat([ones*i, updated_coords*interval])")
__global__ void __launch_bounds__(256) unravel_kernel(const int64_t* __restrict__ lin, int64_t M, int Y, int Z,
                                                      long long ax, long long ay, long long az, long long mul,
                                                      int with_batch, long long batch, int64_t* __restrict__ out) {
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  const int64_t l = __ldg(lin + m);
  const long long z = l % Z;
  const long long r = l / Z;
  const long long y = r % Y;
  const long long x = r / Y;
  int64_t* o = out + m * (with_batch ? 4 : 3);
  if (with_batch) *o++ = batch;
  o[0] = (x + ax) * mul;
  o[1] = (y + ay) * mul;
  o[2] = (z + az) * mul;
}

// dense -> sparse:  out[k] = volume[coords[k]]  (gru_fusion.py:256-261)
__global__ void __launch_bounds__(256) dense_gather_kernel(const float* __restrict__ vol, int X, int Y, int Z, int c,
                                                           const int64_t* __restrict__ coords, int64_t K,
                                                           float* __restrict__ out, int* __restrict__ bad) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= K * c) return;
  const int64_t k = t / c;
  const int j = (int)(t - k * c);
  const int64_t lin = lin3(coords, k, X, Y, Z);
  if (lin < 0) {
    if (j == 0) atomicAdd(bad, 1);
    return;
  }
  out[t] = __ldg(vol + lin * c + j);
}

// rows + constant  (update_map: coords + relative_origin, gru_fusion.py:135/145)
__global__ void __launch_bounds__(256) coords_add_kernel(const int64_t* __restrict__ src, int64_t M, long long ax,
                                                         long long ay, long long az, int64_t* __restrict__ dst) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= M * 3) return;
  const int a = (int)(t % 3);
  dst[t] = __ldg(src + t) + (a == 0 ? ax : a == 1 ? ay : az);
}

#define D3M_NEED_DEVICE(what) \
  D3M_REQUIRE(d3m_device_count() > 0, D3M_ERR_NO_DEVICE, what ": no CUDA device (there is no CPU fallback)")

}  // namespace d3m

using namespace d3m;

extern "C" size_t d3m_sparse_to_dense_workspace(int X, int Y, int Z) {
  return align_up((size_t)X * Y * Z * sizeof(int), 256);
}

extern "C" int d3m_sparse_to_dense(const int64_t* locs, int64_t M, const float* values, float scalar_value, int c,
                                   float default_val, int X, int Y, int Z, float* dense, int* bad_rows,
                                   void* workspace, size_t workspace_bytes, void* stream_) {
  D3M_NEED_DEVICE("d3m_sparse_to_dense");
  D3M_REQUIRE(M >= 0 && M < (1ll << 31) && c >= 1 && X >= 0 && Y >= 0 && Z >= 0 && bad_rows, D3M_ERR_ARG,
              "d3m_sparse_to_dense: bad arguments");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  D3M_CUDA_CHECK(cudaMemsetAsync(bad_rows, 0, sizeof(int), stream));
  const int64_t n_vox = (int64_t)X * Y * Z;
  if (n_vox == 0) return D3M_OK;
  D3M_REQUIRE(dense && aligned16(dense), D3M_ERR_ALIGN, "d3m_sparse_to_dense: dense must be 16-byte aligned");
  {
    LaunchScope ls("s2d_fill", stream);
    fill_kernel<<<nblocks((n_vox * c + 3) / 4, 256), 256, 0, stream>>>(dense, n_vox * c, default_val);
    D3M_CUDA_CHECK(cudaGetLastError());
  }
  if (M == 0) return D3M_OK;
  D3M_REQUIRE(locs != nullptr, D3M_ERR_ARG, "d3m_sparse_to_dense: locs is NULL");
  int* owner = nullptr;
  if (workspace) {  // duplicate-safe ("last row wins") mode
    D3M_REQUIRE(workspace_bytes >= d3m_sparse_to_dense_workspace(X, Y, Z), D3M_ERR_WORKSPACE,
                "d3m_sparse_to_dense: workspace too small");
    owner = static_cast<int*>(workspace);
    D3M_CUDA_CHECK(cudaMemsetAsync(owner, 0xff, (size_t)n_vox * sizeof(int), stream));
    LaunchScope ls("s2d_owner", stream);
    scatter_owner_kernel<<<nblocks(M, 256), 256, 0, stream>>>(locs, M, X, Y, Z, owner, bad_rows);
    D3M_CUDA_CHECK(cudaGetLastError());
  }
  LaunchScope ls("s2d_scatter", stream);
  scatter_rows_kernel<<<nblocks(M * c, 256), 256, 0, stream>>>(locs, M, X, Y, Z, owner, values, values != nullptr, c,
                                                              scalar_value, dense);
  D3M_CUDA_CHECK(cudaGetLastError());
  return D3M_OK;
}

extern "C" int d3m_fbv_mask(const int64_t* global_coords, int64_t M, const int64_t* relative_origin3_host, int X, int Y,
                            int Z, const float* occupied_volume, int64_t* shifted, uint8_t* valid, void* stream_) {
  D3M_NEED_DEVICE("d3m_fbv_mask");
  D3M_REQUIRE(M >= 0 && relative_origin3_host, D3M_ERR_ARG, "d3m_fbv_mask: bad arguments");
  if (M == 0) return D3M_OK;
  D3M_REQUIRE(global_coords && valid, D3M_ERR_ARG, "d3m_fbv_mask: NULL pointer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LaunchScope ls("fbv_mask", stream);
  fbv_mask_kernel<<<nblocks(M, 256), 256, 0, stream>>>(global_coords, M, relative_origin3_host[0],
                                                      relative_origin3_host[1], relative_origin3_host[2], X, Y, Z,
                                                      occupied_volume, shifted, valid);
  D3M_CUDA_CHECK(cudaGetLastError());
  return D3M_OK;
}

extern "C" int d3m_dense_union_flags(const float* vol_a, const float* vol_b, int64_t n_vox, int c, int mode,
                                     uint8_t* flags, void* stream_) {
  D3M_NEED_DEVICE("d3m_dense_union_flags");
  D3M_REQUIRE(n_vox >= 0 && c >= 1 && (mode == 0 || mode == 1), D3M_ERR_ARG, "d3m_dense_union_flags: bad arguments");
  if (n_vox == 0) return D3M_OK;
  D3M_REQUIRE(vol_a && flags, D3M_ERR_ARG, "d3m_dense_union_flags: NULL pointer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LaunchScope ls("dense_union_flags", stream);
  const uintptr_t al = reinterpret_cast<uintptr_t>(vol_a) | reinterpret_cast<uintptr_t>(vol_b);
  if (c > 1 && c % 4 == 0 && c / 4 <= UF_MAX_W && (al & 15u) == 0) {
    union_flags_warp_kernel<float4><<<nblocks(n_vox, 256), 256, 0, stream>>>(
        reinterpret_cast<const float4*>(vol_a), reinterpret_cast<const float4*>(vol_b), n_vox, c / 4, mode, flags);
  } else if (c > 1 && c <= UF_MAX_W) {
    union_flags_warp_kernel<float><<<nblocks(n_vox, 256), 256, 0, stream>>>(vol_a, vol_b, n_vox, c, mode, flags);
  } else {
    union_flags_kernel<<<nblocks(n_vox, 256), 256, 0, stream>>>(vol_a, vol_b, n_vox, c, mode, flags);
  }
  D3M_CUDA_CHECK(cudaGetLastError());
  return D3M_OK;
}

extern "C" int d3m_unravel_coords(const int64_t* linear, int64_t M, int Y, int Z, const int64_t* add3_host,
                                  int64_t multiplier, int with_batch, int64_t batch_index, int64_t* out,
                                  void* stream_) {
  D3M_NEED_DEVICE("d3m_unravel_coords");
  D3M_REQUIRE(M >= 0 && Y >= 1 && Z >= 1, D3M_ERR_ARG, "d3m_unravel_coords: bad arguments");
  if (M == 0) return D3M_OK;
  D3M_REQUIRE(linear && out, D3M_ERR_ARG, "d3m_unravel_coords: NULL pointer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const long long ax = add3_host ? add3_host[0] : 0, ay = add3_host ? add3_host[1] : 0, az = add3_host ? add3_host[2] : 0;
  LaunchScope ls("unravel_coords", stream);
  unravel_kernel<<<nblocks(M, 256), 256, 0, stream>>>(linear, M, Y, Z, ax, ay, az, multiplier, with_batch, batch_index,
                                                     out);
  D3M_CUDA_CHECK(cudaGetLastError());
  return D3M_OK;
}

extern "C" int d3m_dense_gather(const float* volume, int X, int Y, int Z, int c, const int64_t* coords, int64_t K,
                                float* out, int* bad_rows, void* stream_) {
  D3M_NEED_DEVICE("d3m_dense_gather");
  D3M_REQUIRE(K >= 0 && c >= 1 && X >= 0 && Y >= 0 && Z >= 0 && bad_rows, D3M_ERR_ARG, "d3m_dense_gather: bad arguments");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  D3M_CUDA_CHECK(cudaMemsetAsync(bad_rows, 0, sizeof(int), stream));
  if (K == 0) return D3M_OK;
  D3M_REQUIRE(volume && coords && out, D3M_ERR_ARG, "d3m_dense_gather: NULL pointer");
  LaunchScope ls("dense_gather", stream);
  dense_gather_kernel<<<nblocks(K * c, 256), 256, 0, stream>>>(volume, X, Y, Z, c, coords, K, out, bad_rows);
  D3M_CUDA_CHECK(cudaGetLastError());
  return D3M_OK;
}

extern "C" int d3m_coords_add(const int64_t* src, int64_t M, const int64_t* add3_host, int64_t* dst, void* stream_) {
  D3M_NEED_DEVICE("d3m_coords_add");
  D3M_REQUIRE(M >= 0 && add3_host, D3M_ERR_ARG, "d3m_coords_add: bad arguments");
  if (M == 0) return D3M_OK;
  D3M_REQUIRE(src && dst, D3M_ERR_ARG, "d3m_coords_add: NULL pointer");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  LaunchScope ls("coords_add", stream);
  coords_add_kernel<<<nblocks(M * 3, 256), 256, 0, stream>>>(src, M, add3_host[0], add3_host[1], add3_host[2], dst);
  D3M_CUDA_CHECK(cudaGetLastError());
  return D3M_OK;
}
