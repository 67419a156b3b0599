// Marching cubes over a dense (X,Y,Z) float32 volume for sm_100a -- the mesh-export step of the TSDF path
// (reference: skimage.measure.marching_cubes[_lewiner](tsdf_vol, level=0) in core/tsdf/tsdf_volume.py:315,335 and
// core/utils/neucon_utils.py:177; scikit-image is not available here, see mc_tables.py for how the case table is derived).
//
//   pass 1  d3m_mc_flags   one thread per voxel: flags the (up to 3) grid edges leaving the voxel in +x/+y/+z that cross
//                          the level, and for the cube whose low corner the voxel is, one flag per triangle slot
//   (host)  the caller compacts both flag arrays with d3m_compact (ordered: vertices come out sorted by (voxel, axis),
//           triangles by (cube, slot) -- a pure function of the volume)
//   pass 2  d3m_mc_emit    one thread per vertex: position by linear interpolation along its edge, normal = normalised
//                          interpolated central-difference gradient (points towards larger values: for a TSDF into free
//                          space), and the vertex id scattered into a dense edge -> vertex map; then one thread per
//                          triangle: case table -> three cube edges -> three vertex ids.
// "inside" = value < level, exactly as in mc_tables.py; NaN compares false, i.e. counts as outside.
#include "d3m_common.cuh"

namespace d3m {

#include "mc_tables.inc"

struct McGrid {
  const float* vol;
  int X, Y, Z;
  float level;
};

__device__ __forceinline__ float mc_at(const McGrid& g, int x, int y, int z) {
  return __ldg(g.vol + ((int64_t)x * g.Y + y) * g.Z + z);
}

__device__ __forceinline__ int mc_case(const McGrid& g, int x, int y, int z) {
  int c = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (mc_at(g, x + (k & 1), y + ((k >> 1) & 1), z + (k >> 2)) < g.level) c |= 1 << k;
  return c;
}

__global__ void __launch_bounds__(256) mc_flags_kernel(const McGrid g, uint8_t* __restrict__ edge_flags,
                                                       uint8_t* __restrict__ tri_flags) {
  const int64_t nvox = (int64_t)g.X * g.Y * g.Z;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < nvox; i += (int64_t)gridDim.x * 256) {
    const int z = (int)(i % g.Z), y = (int)((i / g.Z) % g.Y), x = (int)(i / ((int64_t)g.Z * g.Y));
    const bool in0 = mc_at(g, x, y, z) < g.level;
    edge_flags[3 * i + 0] = (x + 1 < g.X) && (in0 != (mc_at(g, x + 1, y, z) < g.level));
    edge_flags[3 * i + 1] = (y + 1 < g.Y) && (in0 != (mc_at(g, x, y + 1, z) < g.level));
    edge_flags[3 * i + 2] = (z + 1 < g.Z) && (in0 != (mc_at(g, x, y, z + 1) < g.level));
    int nt = 0;
    if (x + 1 < g.X && y + 1 < g.Y && z + 1 < g.Z) nt = c_mc_ntri[mc_case(g, x, y, z)];
#pragma unroll
    for (int k = 0; k < kMcMaxTri; ++k) tri_flags[kMcMaxTri * i + k] = k < nt;
  }
}

// central difference, one-sided at the border
__device__ __forceinline__ void mc_gradient(const McGrid& g, int x, int y, int z, float grad[3]) {
  const int xm = max(x - 1, 0), xp = min(x + 1, g.X - 1);
  const int ym = max(y - 1, 0), yp = min(y + 1, g.Y - 1);
  const int zm = max(z - 1, 0), zp = min(z + 1, g.Z - 1);
  grad[0] = __fdiv_rn(__fsub_rn(mc_at(g, xp, y, z), mc_at(g, xm, y, z)), (float)max(xp - xm, 1));
  grad[1] = __fdiv_rn(__fsub_rn(mc_at(g, x, yp, z), mc_at(g, x, ym, z)), (float)max(yp - ym, 1));
  grad[2] = __fdiv_rn(__fsub_rn(mc_at(g, x, y, zp), mc_at(g, x, y, zm)), (float)max(zp - zm, 1));
}

__global__ void __launch_bounds__(256) mc_verts_kernel(const McGrid g, const int64_t* __restrict__ edge_list, int64_t nv,
                                                       int* __restrict__ vid, float* __restrict__ verts,
                                                       float* __restrict__ normals) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < nv; i += (int64_t)gridDim.x * 256) {
    const int64_t e = edge_list[i];
    const int axis = (int)(e % 3);
    const int64_t v = e / 3;
    const int z = (int)(v % g.Z), y = (int)((v / g.Z) % g.Y), x = (int)(v / ((int64_t)g.Z * g.Y));
    const int x1 = x + (axis == 0), y1 = y + (axis == 1), z1 = z + (axis == 2);
    const float a = mc_at(g, x, y, z), b = mc_at(g, x1, y1, z1);
    const float t = __fdiv_rn(__fsub_rn(g.level, a), __fsub_rn(b, a));   // in [0, 1]: the edge crosses the level
    float p[3] = {(float)x, (float)y, (float)z};
    p[axis] = __fadd_rn(p[axis], t);
    float g0[3], g1[3], n[3];
    mc_gradient(g, x, y, z, g0);
    mc_gradient(g, x1, y1, z1, g1);
#pragma unroll
    for (int k = 0; k < 3; ++k) n[k] = __fmaf_rn(t, __fsub_rn(g1[k], g0[k]), g0[k]);
    const float len = sqrtf(__fmaf_rn(n[2], n[2], __fmaf_rn(n[1], n[1], __fmul_rn(n[0], n[0]))));
    const float inv = len > 0.0f ? __fdiv_rn(1.0f, len) : 0.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      verts[3 * i + k] = p[k];
      normals[3 * i + k] = __fmul_rn(n[k], inv);
    }
    vid[e] = (int)i;
  }
}

__global__ void __launch_bounds__(256) mc_faces_kernel(const McGrid g, const int64_t* __restrict__ tri_list, int64_t nf,
                                                       const int* __restrict__ vid, int* __restrict__ faces) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < nf; i += (int64_t)gridDim.x * 256) {
    const int64_t s = tri_list[i];
    const int k = (int)(s % kMcMaxTri);
    const int64_t v = s / kMcMaxTri;
    const int z = (int)(v % g.Z), y = (int)((v / g.Z) % g.Y), x = (int)(v / ((int64_t)g.Z * g.Y));
    const int c = mc_case(g, x, y, z);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int e = c_mc_tri[c][3 * k + j];            // cube edge: 4 * axis + bit(first other axis) + 2 * bit(second)
      const int axis = e >> 2, b0 = e & 1, b1 = (e >> 1) & 1;
      int off[3] = {0, 0, 0};
      off[axis == 0 ? 1 : 0] = b0;                      // other axes in ascending order
      off[axis == 2 ? 1 : 2] = b1;
      const int64_t vv = ((int64_t)(x + off[0]) * g.Y + (y + off[1])) * g.Z + (z + off[2]);
      faces[3 * i + j] = vid[3 * vv + axis];
    }
  }
}

static unsigned mc_ctas(int64_t n) {
  int64_t c = (n + 255) / 256;
  if (c > 148 * 16) c = 148 * 16;
  return (unsigned)(c < 1 ? 1 : c);
}

}  // namespace d3m

using namespace d3m;

extern "C" int d3m_mc_max_triangles_per_cube(void) { return kMcMaxTri; }

extern "C" int d3m_mc_flags(const float* volume, int X, int Y, int Z, float level, uint8_t* edge_flags, uint8_t* tri_flags,
                            void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  D3M_REQUIRE(d3m_device_count() > 0, D3M_ERR_NO_DEVICE, "marching cubes: no CUDA device (there is no CPU fallback)");
  D3M_REQUIRE(X >= 1 && Y >= 1 && Z >= 1, D3M_ERR_ARG, "marching cubes: bad volume size");
  D3M_REQUIRE(volume && edge_flags && tri_flags, D3M_ERR_ARG, "marching cubes: NULL pointer");
  McGrid g;
  g.vol = volume; g.X = X; g.Y = Y; g.Z = Z; g.level = level;
  LaunchScope ls("mc_flags", stream);
  mc_flags_kernel<<<mc_ctas((int64_t)X * Y * Z), 256, 0, stream>>>(g, edge_flags, tri_flags);
  D3M_CUDA_CHECK(cudaGetLastError());
  return D3M_OK;
}

extern "C" int d3m_mc_emit(const float* volume, int X, int Y, int Z, float level, const int64_t* edge_list, int64_t n_verts,
                           const int64_t* tri_list, int64_t n_faces, int* edge_to_vertex, float* verts, float* normals,
                           int* faces, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  D3M_REQUIRE(d3m_device_count() > 0, D3M_ERR_NO_DEVICE, "marching cubes: no CUDA device (there is no CPU fallback)");
  D3M_REQUIRE(X >= 1 && Y >= 1 && Z >= 1 && n_verts >= 0 && n_faces >= 0 && n_verts < (1ll << 31), D3M_ERR_ARG,
              "marching cubes: bad sizes");
  McGrid g;
  g.vol = volume; g.X = X; g.Y = Y; g.Z = Z; g.level = level;
  if (n_verts > 0) {
    D3M_REQUIRE(volume && edge_list && edge_to_vertex && verts && normals, D3M_ERR_ARG, "marching cubes: NULL pointer");
    LaunchScope ls("mc_verts", stream);
    mc_verts_kernel<<<mc_ctas(n_verts), 256, 0, stream>>>(g, edge_list, n_verts, edge_to_vertex, verts, normals);
    D3M_CUDA_CHECK(cudaGetLastError());
  }
  if (n_faces > 0) {
    D3M_REQUIRE(volume && tri_list && edge_to_vertex && faces, D3M_ERR_ARG, "marching cubes: NULL pointer");
    LaunchScope ls("mc_faces", stream);
    mc_faces_kernel<<<mc_ctas(n_faces), 256, 0, stream>>>(g, tri_list, n_faces, edge_to_vertex, faces);
    D3M_CUDA_CHECK(cudaGetLastError());
  }
  return D3M_OK;
}
