// libd3m core: error reporting, device probe, and the NCHW <-> NHWC relayout of the per-view feature maps.
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "d3m_common.cuh"

namespace d3m {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return D3M_ERR_CUDA + (int)e;
}

// D3M_PDL: 0 = never, 1 = always, unset / "auto" = per call (PdlScope): programmatic dependent launch pays where the GPU
// waits for the host between small kernels (an eager fragment-sized step: 0.51 -> 0.44 ms) and costs where kernels run
// back to back -- CUDA-graph replay of the same step 0.235 -> 0.256 ms, 64 batched fragments 10.8 -> 12.3 ms
// (profiles/r02t_*): the dependent's CTAs take whatever SM slots free up first while they wait, and then run from that
// uneven placement.
static int pdl_mode() {
  static const int mode = [] {
    const char* e = getenv("D3M_PDL");
    if (!e || !*e || e[0] == 'a') return 2;
    return e[0] == '0' ? 0 : 1;
  }();
  return mode;
}
static thread_local bool g_pdl_call = true;

bool pdl_enabled() {
  const int m = pdl_mode();
  return m == 2 ? g_pdl_call : m == 1;
}

PdlScope::PdlScope(cudaStream_t stream, long long work_items) : prev_(g_pdl_call) {
  bool on = work_items <= kPdlMaxWorkItems;
  if (on && pdl_mode() == 2) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone) on = false;
  }
  g_pdl_call = on;
}
PdlScope::~PdlScope() { g_pdl_call = prev_; }

int current_device_sms() {
  static std::atomic<int> cache[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 0 && dev < 64) {
    const int c = cache[dev].load(std::memory_order_relaxed);
    if (c > 0) return c;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (dev >= 0 && dev < 64) cache[dev].store(sms, std::memory_order_relaxed);
  return sms;
}

cudaError_t ensure_dynamic_smem(const void* kernel, size_t bytes) {
  // the attribute is per (function, device); remember the largest value set so far
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, size_t> set;
  if (bytes <= 48 * 1024) return cudaSuccess;   // the default limit needs no opt-in
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(mu);
  size_t& cur = set[{kernel, dev}];
  if (bytes <= cur) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) cur = bytes;
  return e;
}

static std::atomic<long long> g_launches{0};
static std::atomic<bool> g_profiling{false};
static std::mutex g_prof_mu;
struct ProfRec { const char* name; cudaEvent_t a, b; };
static std::vector<ProfRec> g_prof;

LaunchScope::LaunchScope(const char* name, cudaStream_t stream) : name_(name), stream_(stream), start_(nullptr) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (g_profiling.load(std::memory_order_relaxed)) {
    if (cudaEventCreate(&start_) == cudaSuccess) cudaEventRecord(start_, stream_);
    else start_ = nullptr;
  }
}

LaunchScope::~LaunchScope() {
  if (!start_) return;
  cudaEvent_t stop = nullptr;
  if (cudaEventCreate(&stop) == cudaSuccess) {
    cudaEventRecord(stop, stream_);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back({name_, start_, stop});
  } else {
    cudaEventDestroy(start_);
  }
}

// Clears `bytes` (multiple of 4, 16-byte aligned start) with a kernel instead of a memset node, so that the clear takes
// part in the programmatic-dependent-launch chain of d3m_common.cuh (a memset between two kernels serialises them fully).
__global__ void __launch_bounds__(256) zero_words_kernel(uint32_t* __restrict__ p, int64_t n_words) {
  pdl_enter();
  const int64_t nv = n_words >> 2;
  uint4* v = reinterpret_cast<uint4*>(p);
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < nv; i += (int64_t)gridDim.x * 256)
    v[i] = make_uint4(0u, 0u, 0u, 0u);
  if (blockIdx.x == 0 && threadIdx.x < (int)(n_words & 3)) p[(nv << 2) + threadIdx.x] = 0u;
}

int zero_async(void* p, size_t bytes, cudaStream_t stream) {
  if (bytes == 0) return D3M_OK;
  if (!pdl_enabled() || (bytes & 3) || !aligned16(p)) {
    D3M_CUDA_CHECK(cudaMemsetAsync(p, 0, bytes, stream));
    return D3M_OK;
  }
  const int64_t words = (int64_t)(bytes >> 2);
  int64_t ctas = ((words >> 2) + 255) / 256;
  if (ctas > 148 * 8) ctas = 148 * 8;
  if (ctas < 1) ctas = 1;
  LaunchScope ls("zero_words", stream);
  D3M_CUDA_CHECK(launch_k(zero_words_kernel, dim3((unsigned)ctas), dim3(256), 0, stream, static_cast<uint32_t*>(p), words));
  return D3M_OK;
}

// (n_maps, A, Bn) -> (n_maps, Bn, A) through a padded 32x32 shared-memory tile: both sides coalesced.
__global__ void __launch_bounds__(256) transpose_maps_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                             int A, int Bn) {
  pdl_enter();
  __shared__ float tile[32][33];
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

static int transpose_maps(const float* src, float* dst, int64_t n_maps, int A, int Bn, cudaStream_t stream) {
  D3M_REQUIRE(d3m_device_count() > 0, D3M_ERR_NO_DEVICE, "relayout: no CUDA device (there is no CPU fallback)");
  D3M_REQUIRE(src && dst && n_maps >= 0 && A >= 1 && Bn >= 1, D3M_ERR_ARG, "relayout: bad arguments");
  if (n_maps == 0) return D3M_OK;
  for (int64_t m0 = 0; m0 < n_maps; m0 += 65535) {
    const unsigned nz = (unsigned)((n_maps - m0) < 65535 ? (n_maps - m0) : 65535);
    dim3 grid((Bn + 31) / 32, (A + 31) / 32, nz);
    LaunchScope ls("relayout_transpose", stream);
    launch_k(transpose_maps_kernel, dim3(grid), dim3(256), 0, stream, src + m0 * (int64_t)A * Bn, dst + m0 * (int64_t)A * Bn, A, Bn);
    D3M_CUDA_CHECK(cudaGetLastError());
  }
  return D3M_OK;
}

}  // namespace d3m

extern "C" int d3m_version(void) { return D3M_VERSION; }

extern "C" const char* d3m_last_error(void) { return d3m::g_err; }

extern "C" int d3m_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" int d3m_current_device(void) {
  int d = -1;
  if (d3m_device_count() <= 0 || cudaGetDevice(&d) != cudaSuccess) {
    cudaGetLastError();
    return -1;
  }
  return d;
}

extern "C" int64_t d3m_kernel_launches(void) { return (int64_t)d3m::g_launches.load(); }

extern "C" int d3m_profile_begin(void) {
  std::lock_guard<std::mutex> lk(d3m::g_prof_mu);
  for (auto& r : d3m::g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  d3m::g_prof.clear();
  d3m::g_profiling.store(true);
  return D3M_OK;
}

extern "C" int d3m_profile_end(char* json_out, size_t cap) {
  d3m::g_profiling.store(false);
  std::lock_guard<std::mutex> lk(d3m::g_prof_mu);
  std::map<std::string, std::pair<long long, double>> agg;
  for (auto& r : d3m::g_prof) {
    float ms = 0.f;
    cudaEventSynchronize(r.b);
    if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
      auto& e = agg[r.name];
      e.first += 1;
      e.second += ms;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  d3m::g_prof.clear();
  std::string js = "{";
  bool first = true;
  for (auto& kv : agg) {
    char buf[256];
    snprintf(buf, sizeof(buf), "%s\"%s\": {\"n\": %lld, \"ms\": %.6f}", first ? "" : ", ", kv.first.c_str(),
             kv.second.first, kv.second.second);
    js += buf;
    first = false;
  }
  js += "}";
  D3M_REQUIRE(json_out && cap > js.size(), D3M_ERR_ARG, "profile_end: buffer too small (%zu needed)", js.size() + 1);
  memcpy(json_out, js.c_str(), js.size() + 1);
  return D3M_OK;
}

extern "C" int d3m_feats_nchw_to_nhwc(const float* src, float* dst, int64_t n_maps, int C, int H, int W, void* stream) {
  return d3m::transpose_maps(src, dst, n_maps, C, H * W, static_cast<cudaStream_t>(stream));
}

extern "C" int d3m_feats_nhwc_to_nchw(const float* src, float* dst, int64_t n_maps, int C, int H, int W, void* stream) {
  return d3m::transpose_maps(src, dst, n_maps, H * W, C, static_cast<cudaStream_t>(stream));
}

// ---- peer-visible device memory (one process per GPU, one box): CUDA IPC -------------------------------------------
// The exchange buffers of the voxel-sharded backward live in plain cudaMalloc memory owned by this library; the 64-byte
// IPC handle travels through the caller's process group (torch.distributed.all_gather_object) and every peer maps the
// buffer into its own address space.  Kernels then store to it directly over NVLink.
extern "C" int d3m_p2p_alloc(size_t bytes, void** dev_ptr, unsigned char* handle64_host) {
  D3M_REQUIRE(d3m_device_count() > 0, D3M_ERR_NO_DEVICE, "d3m_p2p_alloc: no CUDA device");
  D3M_REQUIRE(bytes > 0 && dev_ptr && handle64_host, D3M_ERR_ARG, "d3m_p2p_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  D3M_CUDA_CHECK(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  const cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) { cudaFree(p); return d3m::cuda_fail(e, "cudaIpcGetMemHandle"); }
  memcpy(handle64_host, &h, 64);
  *dev_ptr = p;
  return D3M_OK;
}

extern "C" int d3m_p2p_open(const unsigned char* handle64_host, void** dev_ptr) {
  D3M_REQUIRE(d3m_device_count() > 0, D3M_ERR_NO_DEVICE, "d3m_p2p_open: no CUDA device");
  D3M_REQUIRE(handle64_host && dev_ptr, D3M_ERR_ARG, "d3m_p2p_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64_host, 64);
  D3M_CUDA_CHECK(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return D3M_OK;
}

extern "C" int d3m_p2p_close(void* dev_ptr) {
  if (!dev_ptr) return D3M_OK;
  D3M_CUDA_CHECK(cudaIpcCloseMemHandle(dev_ptr));
  return D3M_OK;
}

extern "C" int d3m_p2p_free(void* dev_ptr) {
  if (!dev_ptr) return D3M_OK;
  D3M_CUDA_CHECK(cudaFree(dev_ptr));
  return D3M_OK;
}

// ---- all-gather of per-voxel rows as peer stores --------------------------------------------------------------------
// Every rank writes its n_local rows (row_words 32-bit words each) into EVERY rank's full-size buffer at the rows' global
// positions: contiguous ranges (block == 0: global = begin + i) or block-cyclic ranges (global = ((i / block) * world +
// rank) * block + i % block, shard.voxel_blocks).  The next collective of the caller on the same streams (the all-reduce
// of the depth sums) is the barrier after which every buffer is complete: no gather collective, no padding, no
// re-ordering pass (the rows land in the scene's voxel order).
namespace d3m {
__global__ void __launch_bounds__(256) p2p_scatter_rows_kernel(const uint32_t* __restrict__ src, int64_t n_local, int row_words,
                                                               int64_t begin, int64_t block, int world, int rank,
                                                               uint32_t* const* __restrict__ peers) {
  pdl_enter();
  const int64_t total = n_local * row_words;
  for (int64_t k = (int64_t)blockIdx.x * 256 + threadIdx.x; k < total; k += (int64_t)gridDim.x * 256) {
    const int64_t i = k / row_words;
    const int w = (int)(k - i * row_words);
    const int64_t g = block > 0 ? ((i / block) * world + rank) * block + i % block : begin + i;
    const uint32_t v = __ldg(src + k);
    for (int r = 0; r < world; ++r) peers[r][g * row_words + w] = v;
  }
}

// one-word rows whose runs are 16-byte aligned on both sides (block-cyclic with block % 4 == 0, or begin % 4 == 0): four
// rows per store -- peer stores of 16 bytes instead of 4
__global__ void __launch_bounds__(256) p2p_scatter_rows4_kernel(const uint4* __restrict__ src, int64_t n_quads, int64_t begin,
                                                                int64_t block, int world, int rank,
                                                                uint32_t* const* __restrict__ peers) {
  pdl_enter();
  for (int64_t q = (int64_t)blockIdx.x * 256 + threadIdx.x; q < n_quads; q += (int64_t)gridDim.x * 256) {
    const int64_t i = q * 4;
    const int64_t g = block > 0 ? ((i / block) * world + rank) * block + i % block : begin + i;
    const uint4 v = __ldg(src + q);
    for (int r = 0; r < world; ++r) *reinterpret_cast<uint4*>(peers[r] + g) = v;
  }
}

// ---- all-reduce of a few doubles + barrier across the ranks of one box, through peer memory --------------------------
// Every rank owns a mailbox of `world` slots {flag, payload[kSyncPayload]} in IPC-shared memory.  One CTA per rank:
// thread r writes this rank's payload into slot `rank` of rank r's mailbox, fences at system scope and raises the slot's
// flag to `epoch` (the caller's call counter, identical on every rank); then thread r waits until slot r of its OWN
// mailbox shows `epoch`, and the payloads are added in ascending rank order (deterministic).  Peer stores issued by earlier
// kernels of the same stream (count rows, gradient slots) were performed before this kernel started, so once every flag
// has arrived those stores have arrived too: the same launch is the barrier of the fused exchanges.  ~5 us against
// ~25 us for a 4-byte NCCL all-reduce on 8 GPUs.
constexpr int kSyncPayload = 192;   // doubles (3 per fragment, up to 64 fragments)
struct SyncSlot {
  unsigned long long flag;
  unsigned long long pad;
  double payload[2][kSyncPayload];   // by epoch parity: a rank may be one call ahead of the slowest, never two
};
__global__ void __launch_bounds__(64) p2p_sync_kernel(SyncSlot* const* __restrict__ mailboxes, int world, int rank,
                                                      unsigned long long epoch, const double* __restrict__ payload, int n,
                                                      double* __restrict__ out) {
  pdl_enter();
  const int r = threadIdx.x;
  if (r < world) {
    SyncSlot* dst = mailboxes[r] + rank;
    for (int i = 0; i < n; ++i) dst->payload[epoch & 1ull][i] = payload[i];
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long*>(&dst->flag) = epoch;
    const volatile unsigned long long* mine = &mailboxes[rank][r].flag;
    const long long t0 = clock64();
    while (*mine < epoch) {
      if (clock64() - t0 > 40000000000ll) __trap();   // ~20 s: a peer never arrived
      __nanosleep(100);
    }
    __threadfence_system();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double a = 0.0;
    for (int q = 0; q < world; ++q)
      a += *reinterpret_cast<const volatile double*>(&mailboxes[rank][q].payload[epoch & 1ull][i]);
    out[i] = a;
  }
}
}  // namespace d3m

extern "C" int d3m_p2p_scatter_rows(const void* src, int64_t n_local, int row_bytes, int64_t begin, int64_t block,
                                    void* const* peer_dst_dev_table, int world, int rank, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  D3M_REQUIRE(d3m_device_count() > 0, D3M_ERR_NO_DEVICE, "d3m_p2p_scatter_rows: no CUDA device");
  D3M_REQUIRE(n_local >= 0 && row_bytes >= 4 && row_bytes % 4 == 0 && begin >= 0 && block >= 0 && world >= 1 && rank >= 0 &&
                  rank < world && peer_dst_dev_table,
              D3M_ERR_ARG, "d3m_p2p_scatter_rows: bad arguments");
  if (n_local == 0) return D3M_OK;
  D3M_REQUIRE(src, D3M_ERR_ARG, "d3m_p2p_scatter_rows: NULL source");
  d3m::LaunchScope ls("p2p_scatter_rows", stream);
  const bool quads = row_bytes == 4 && n_local % 4 == 0 && d3m::aligned16(src) &&
                     (block > 0 ? block % 4 == 0 : begin % 4 == 0);   // peer buffers are cudaMalloc bases: 256-byte aligned
  if (quads) {
    int64_t ctas = (n_local / 4 + 255) / 256;
    if (ctas > 148 * 8) ctas = 148 * 8;
    d3m::launch_k(d3m::p2p_scatter_rows4_kernel, dim3((unsigned)ctas), dim3(256), 0, stream, static_cast<const uint4*>(src),
                  n_local / 4, begin, block, world, rank, reinterpret_cast<uint32_t* const*>(peer_dst_dev_table));
  } else {
    const int64_t total = n_local * (row_bytes / 4);
    int64_t ctas = (total + 255) / 256;
    if (ctas > 148 * 8) ctas = 148 * 8;
    d3m::launch_k(d3m::p2p_scatter_rows_kernel, dim3((unsigned)ctas), dim3(256), 0, stream, static_cast<const uint32_t*>(src),
                  n_local, row_bytes / 4, begin, block, world, rank, reinterpret_cast<uint32_t* const*>(peer_dst_dev_table));
  }
  D3M_CUDA_CHECK(cudaGetLastError());
  return D3M_OK;
}

extern "C" size_t d3m_p2p_sync_mailbox_bytes(int world) { return world > 0 ? sizeof(d3m::SyncSlot) * (size_t)world : 0; }

extern "C" int d3m_p2p_sync(void* const* mailbox_dev_table, int world, int rank, unsigned long long epoch,
                            const double* payload, int n, double* out, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  D3M_REQUIRE(d3m_device_count() > 0, D3M_ERR_NO_DEVICE, "d3m_p2p_sync: no CUDA device");
  D3M_REQUIRE(mailbox_dev_table && world >= 1 && world <= 64 && rank >= 0 && rank < world && epoch > 0 && n >= 0 &&
                  n <= d3m::kSyncPayload && (n == 0 || (payload && out)),
              D3M_ERR_ARG, "d3m_p2p_sync: bad arguments (at most 64 ranks, %d doubles)", d3m::kSyncPayload);
  d3m::LaunchScope ls("p2p_sync", stream);
  d3m::launch_k(d3m::p2p_sync_kernel, dim3(1), dim3(64), 0, stream,
                reinterpret_cast<d3m::SyncSlot* const*>(mailbox_dev_table), world, rank, epoch, payload, n, out);
  D3M_CUDA_CHECK(cudaGetLastError());
  return D3M_OK;
}
