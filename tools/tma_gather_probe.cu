// Load-path experiment for the bilinear gather of back_project: the same sample list -- 4 corner
// texels of C floats per sample, channels-last maps resident in L2, 5 samples per warp round like bp_fwd_kernel<*,6,1> --
// fetched three ways, each feeding the same FMA chain and writing the same (S, C) result:
//
//   ldg      4 x LDG.E.128 per lane (6 lanes x float4 = 24 channels), the shipped path
//   gather4  ONE cp.async.bulk.tensor.2d.tile::gather4 per sample (rows = the 4 corner texels of a [texel, C] tensor map,
//            384 B), issued by lane 0 into a per-warp shared-memory ring (DEPTH stages of 5 samples), mbarrier completion,
//            the lane groups read their corners back with LDS.128
//   bulk     TWO cp.async.bulk (1-D, 2*C*4 = 192 B: the texel pair x0, x0+1 of one image row) per sample, same ring
//
// Build + run on the GPU box:   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o /tmp/tma_probe
//                                    tools/tma_gather_probe.cu -lcuda && /tmp/tma_probe
// Prints one line per variant: us per launch (best of 20, maps L2-warm), samples/s, GB/s of corner texels, and whether the
// result is bit-identical to the ldg variant.  Not part of libd3m.so.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));   \
      exit(1);                                                                     \
    }                                                                              \
  } while (0)

constexpr int C = 24, C4 = C / 4;       // level-2 channel count: 96-byte texels
constexpr int G = 6;                    // lanes per sample
constexpr int NS = 32 / G;              // samples per warp round (5)
constexpr int WARPS = 4;
constexpr int DEPTH = 4;                // ring stages per warp
constexpr int SLOT = 4 * C * 4;         // bytes per sample in the ring: 4 corners (384 B = 3 x 128)

struct SampleRec { int t00; float fx, fy; };   // texel index of the (x0, y0) corner; x0+1 / y0+1 always inside the map

__device__ __forceinline__ float chain(float a, float b, float c, float d, float nw, float ne, float sw, float se) {
  return __fmaf_rn(d, se, __fmaf_rn(c, sw, __fmaf_rn(b, ne, __fmul_rn(a, nw))));
}

__device__ __forceinline__ void weights(float fx, float fy, float& nw, float& ne, float& sw, float& se) {
  const float wx0 = __fsub_rn(1.0f, fx), wy0 = __fsub_rn(1.0f, fy);
  nw = __fmul_rn(wx0, wy0); ne = __fmul_rn(fx, wy0); sw = __fmul_rn(wx0, fy); se = __fmul_rn(fx, fy);
}

// ---- variant 1: LDG.E.128 -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(WARPS * 32) gather_ldg(const float4* __restrict__ maps, const SampleRec* __restrict__ rec,
                                                         int64_t S, int W, float4* __restrict__ out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane / G, gl = lane % G;
  const int64_t rounds = (S + NS - 1) / NS;
  for (int64_t r = (int64_t)blockIdx.x * WARPS + warp; r < rounds; r += (int64_t)gridDim.x * WARPS) {
    const int64_t s = r * NS + g;
    if (g >= NS || s >= S) continue;
    const SampleRec q = rec[s];
    float nw, ne, sw, se;
    weights(q.fx, q.fy, nw, ne, sw, se);
    const float4* b = maps + (int64_t)q.t00 * C4 + gl;
    const float4 a00 = __ldg(b), a01 = __ldg(b + C4), a10 = __ldg(b + (int64_t)W * C4), a11 = __ldg(b + (int64_t)(W + 1) * C4);
    float4 o;
    o.x = chain(a00.x, a01.x, a10.x, a11.x, nw, ne, sw, se);
    o.y = chain(a00.y, a01.y, a10.y, a11.y, nw, ne, sw, se);
    o.z = chain(a00.z, a01.z, a10.z, a11.z, nw, ne, sw, se);
    o.w = chain(a00.w, a01.w, a10.w, a11.w, nw, ne, sw, se);
    out[s * C4 + gl] = o;
  }
}

// ---- shared-memory ring helpers ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* m, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(m)), "r"(count));
}
__device__ __forceinline__ void mbar_expect(uint64_t* m, int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(m)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* m, int parity) {
  asm volatile(
      "{\n.reg .pred p;\nW_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra D_%=;\nbra W_%=;\nD_%=:\n}" ::"r"(smem_u32(m)), "r"(parity) : "memory");
}

// MODE 0: tile::gather4 through a tensor map; MODE 1: two 1-D bulk copies of a texel pair
template <int MODE>
__global__ void __launch_bounds__(WARPS * 32) gather_tma(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ maps,
                                                         const SampleRec* __restrict__ rec, int64_t S, int W,
                                                         float4* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[WARPS][DEPTH];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane / G, gl = lane % G;
  unsigned char* ring = smem + (size_t)warp * DEPTH * NS * SLOT;
  if (lane == 0)
    for (int d = 0; d < DEPTH; ++d) mbar_init(&bars[warp][d], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const int64_t rounds = (S + NS - 1) / NS;
  const int64_t r0 = (int64_t)blockIdx.x * WARPS + warp, stride = (int64_t)gridDim.x * WARPS;

  auto issue = [&](int64_t r, int stage) {   // lane 0: the copies of round r into ring stage `stage`
    const int64_t s0 = r * NS;
    const int n = (int)((S - s0) < NS ? (S - s0) : NS);
    mbar_expect(&bars[warp][stage], n * SLOT);
    for (int i = 0; i < n; ++i) {
      const int t = rec[s0 + i].t00;
      const unsigned dst = smem_u32(ring + ((size_t)stage * NS + i) * SLOT);
      const unsigned mb = smem_u32(&bars[warp][stage]);
      if (MODE == 0) {
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
            " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst), "l"(&tmap), "r"(mb), "r"(0), "r"(t), "r"(t + 1),
            "r"(t + W), "r"(t + W + 1)
            : "memory");
      } else {
        const float* src = maps + (int64_t)t * C;
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                     "l"(src), "r"(2 * C * 4), "r"(mb)
                     : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         dst + 2 * C * 4),
                     "l"(src + (int64_t)W * C), "r"(2 * C * 4), "r"(mb)
                     : "memory");
      }
    }
  };

  if (lane == 0)
    for (int d = 0; d < DEPTH; ++d)
      if (r0 + d * stride < rounds) issue(r0 + d * stride, d);
  int it = 0;
  for (int64_t r = r0; r < rounds; r += stride, ++it) {
    const int stage = it % DEPTH, parity = (it / DEPTH) & 1;
    mbar_wait(&bars[warp][stage], parity);
    const int64_t s = r * NS + g;
    if (g < NS && s < S) {
      const SampleRec q = rec[s];
      float nw, ne, sw, se;
      weights(q.fx, q.fy, nw, ne, sw, se);
      const float4* b = reinterpret_cast<const float4*>(ring + ((size_t)stage * NS + g) * SLOT) + gl;
      const float4 a00 = b[0], a01 = b[C4], a10 = b[2 * C4], a11 = b[3 * C4];
      float4 o;
      o.x = chain(a00.x, a01.x, a10.x, a11.x, nw, ne, sw, se);
      o.y = chain(a00.y, a01.y, a10.y, a11.y, nw, ne, sw, se);
      o.z = chain(a00.z, a01.z, a10.z, a11.z, nw, ne, sw, se);
      o.w = chain(a00.w, a01.w, a10.w, a11.w, nw, ne, sw, se);
      out[s * C4 + gl] = o;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads of the stage before its async refill
    __syncwarp();
    const int64_t rn = r + (int64_t)DEPTH * stride;
    if (lane == 0 && rn < rounds) issue(rn, stage);
  }
}

int main(int argc, char** argv) {
  const int V = 9, H = 120, W = 160;
  const int64_t S = argc > 1 ? atoll(argv[1]) : 688661;   // valid samples of the level-2 fragment pass
  const int64_t texels = (int64_t)V * H * W;
  std::vector<float> h_maps((size_t)texels * C);
  uint64_t x = 88172645463325252ull;
  auto rnd = [&]() { x ^= x << 13; x ^= x >> 7; x ^= x << 17; return x; };
  for (auto& f : h_maps) f = (float)((int64_t)(rnd() % 2001) - 1000) * 1e-3f;
  std::vector<SampleRec> h_rec((size_t)S);
  for (int64_t s = 0; s < S; s += NS) {   // a round = 5 neighbouring voxels: the same view, adjacent pixels
    const int v = (int)(rnd() % V), by = (int)(rnd() % (H - 1)), bx = (int)(rnd() % (W - 1 - NS));
    for (int i = 0; i < NS && s + i < S; ++i) {
      SampleRec q;
      q.t00 = (v * H + by) * W + bx + i;
      q.fx = (float)(rnd() % 1000) * 1e-3f;
      q.fy = (float)(rnd() % 1000) * 1e-3f;
      h_rec[(size_t)(s + i)] = q;
    }
  }
  float *d_maps, *d_out[3];
  SampleRec* d_rec;
  CK(cudaMalloc(&d_maps, h_maps.size() * 4));
  CK(cudaMalloc(&d_rec, h_rec.size() * sizeof(SampleRec)));
  for (auto& p : d_out) { CK(cudaMalloc(&p, (size_t)S * C * 4)); CK(cudaMemset(p, 0, (size_t)S * C * 4)); }
  CK(cudaMemcpy(d_maps, h_maps.data(), h_maps.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_rec, h_rec.data(), h_rec.size() * sizeof(SampleRec), cudaMemcpyHostToDevice));

  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  {
    const cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)texels};
    const cuuint64_t strides[1] = {(cuuint64_t)C * 4};
    const cuuint32_t box[2] = {(cuuint32_t)C, 1};
    const cuuint32_t estr[2] = {1, 1};
    CUresult rc = cuTensorMapEncodeTiled(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_maps, dims, strides, box, estr,
                                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                         CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { fprintf(stderr, "cuTensorMapEncodeTiled failed: %d\n", (int)rc); return 1; }
  }
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const size_t ring_bytes = (size_t)WARPS * DEPTH * NS * SLOT;
  CK(cudaFuncSetAttribute(gather_tma<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_bytes));
  CK(cudaFuncSetAttribute(gather_tma<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_bytes));
  const int64_t rounds = (S + NS - 1) / NS;
  const char* names[3] = {"ldg    ", "gather4", "bulk   "};
  std::vector<float> h_ref((size_t)S * C), h_out((size_t)S * C);
  printf("# S = %lld samples, C = %d (%d-byte texels), maps %.1f MB (L2-resident), %d SMs, ring %zu B per CTA\n",
         (long long)S, C, C * 4, h_maps.size() * 4 / 1e6, sms, ring_bytes);
  for (int ctas_per_sm : {6, 8, 12}) {
    for (int k = 0; k < 3; ++k) {
      int64_t grid = (int64_t)sms * ctas_per_sm;
      if (grid * WARPS > rounds) grid = (rounds + WARPS - 1) / WARPS;
      auto launch = [&]() {
        if (k == 0) gather_ldg<<<(unsigned)grid, WARPS * 32>>>((const float4*)d_maps, d_rec, S, W, (float4*)d_out[0]);
        else if (k == 1) gather_tma<0><<<(unsigned)grid, WARPS * 32, ring_bytes>>>(tmap, d_maps, d_rec, S, W, (float4*)d_out[1]);
        else gather_tma<1><<<(unsigned)grid, WARPS * 32, ring_bytes>>>(tmap, d_maps, d_rec, S, W, (float4*)d_out[2]);
      };
      for (int i = 0; i < 3; ++i) launch();
      CK(cudaGetLastError());
      CK(cudaDeviceSynchronize());
      cudaEvent_t a, b;
      CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
      float best = 1e30f;
      for (int i = 0; i < 20; ++i) {
        CK(cudaEventRecord(a)); launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
      }
      CK(cudaMemcpy(k == 0 ? h_ref.data() : h_out.data(), d_out[k], (size_t)S * C * 4, cudaMemcpyDeviceToHost));
      const bool same = k == 0 || memcmp(h_ref.data(), h_out.data(), (size_t)S * C * 4) == 0;
      printf("%s  %2d CTAs/SM  %8.2f us  %7.2f G samples/s  %8.1f GB/s corner texels  bit-identical to ldg: %s\n", names[k],
             ctas_per_sm, best * 1e3, S / (best * 1e-3) / 1e9, (double)S * 4 * C * 4 / (best * 1e-3) / 1e9, same ? "yes" : "NO");
    }
  }
  return 0;
}
