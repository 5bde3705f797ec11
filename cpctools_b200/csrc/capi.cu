// capi.cu -- error plumbing, device queries and the synthetic-data generator of libsoapb200.
#include "common.cuh"

#include <stdarg.h>

#include <mutex>
#include <vector>

namespace soapb {

std::atomic<int64_t> g_launches{0};

std::string &last_error() {
  static thread_local std::string msg;
  return msg;
}

int set_error(int code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  last_error() = buf;
  return code;
}

int check_launch(const char *what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(SOAP_ECUDA, "%s launch failed: %s", what, cudaGetErrorString(e));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return SOAP_OK;
}

static int query_attr(cudaDeviceAttr attr, int dflt) {
  int dev = 0, val = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return dflt;
  if (cudaDeviceGetAttribute(&val, attr, dev) != cudaSuccess) return dflt;
  return val;
}
int num_sms() { return query_attr(cudaDevAttrMultiProcessorCount, 148); }
static std::atomic<int> g_reserved_sms{0};
// SMs the persistent kernels leave free (a collective or a copy kernel running beside them needs somewhere to go)
int persistent_sms() {
  const int n = num_sms() - g_reserved_sms.load(std::memory_order_relaxed);
  return n < 1 ? 1 : n;
}
int max_smem_optin() { return query_attr(cudaDevAttrMaxSharedMemoryPerBlockOptin, 232448); }

// ---- profiling hook ------------------------------------------------------------------------
struct ProfRec {
  int id;
  cudaEvent_t start, stop;
};
static std::mutex g_prof_mu;
static std::vector<ProfRec *> g_prof;
static std::atomic<int> g_prof_on{0};

ProfScope::ProfScope(int kernel_id, cudaStream_t s) : id(kernel_id), stream(s), rec(nullptr) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  auto *r = new ProfRec{kernel_id, nullptr, nullptr};
  if (cudaEventCreate(&r->start) != cudaSuccess || cudaEventCreate(&r->stop) != cudaSuccess) {
    delete r;
    return;
  }
  cudaEventRecord(r->start, stream);
  rec = r;
}
ProfScope::~ProfScope() {
  if (!rec) return;
  auto *r = static_cast<ProfRec *>(rec);
  cudaEventRecord(r->stop, stream);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(r);
}

// splitmix64 finaliser on (seed, index): cheap, stateless, reproducible on the host
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

template <typename T>
__global__ void fill_uniform_kernel(T *out, long long n, uint64_t seed) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint64_t r = mix64(seed * 0xD1342543DE82EF95ull + (uint64_t)i);
    if (sizeof(T) == 8)
      out[i] = (T)((double)(r >> 11) * (1.0 / 9007199254740992.0));
    else
      out[i] = (T)((float)(r >> 40) * (1.0f / 16777216.0f));
  }
}

// out[r][g * width + a] = buf[g][r][a]: turns the [shards][rows][width] buffer an all-gather fills into the
// [rows][shards * width] array every rank wants (one pass, 128-bit accesses when the width allows)
__global__ void __launch_bounds__(256) interleave_shards_kernel(const double *__restrict__ buf, double *__restrict__ out,
                                                               long long shards, long long rows, long long width) {
  const long long chunks = (width + 1) / 2;  // pairs of doubles
  const long long total = shards * rows * chunks;
  const bool vec = (width % 2) == 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long c = i % chunks, gr = i / chunks, r = gr % rows, g = gr / rows;
    const double *src = buf + (g * rows + r) * width + 2 * c;
    double *dst = out + (r * shards + g) * width + 2 * c;
    if (vec) {
      const int4 v = ldg_stream16(src);
      stg_stream16(dst, v);
    } else {
      dst[0] = src[0];
      if (2 * c + 1 < width) dst[1] = src[1];
    }
  }
}

}  // namespace soapb

using namespace soapb;

extern "C" int soap_interleave_shards(const double *buf, double *out, int64_t n_shards, int64_t rows, int64_t width,
                                      void *stream) {
  SOAP_REQUIRE(n_shards >= 1 && rows >= 0 && width >= 0, "soap_interleave_shards: bad shape");
  if (rows == 0 || width == 0) return SOAP_OK;
  SOAP_REQUIRE(buf && out, "soap_interleave_shards: null buffer");
  SOAP_REQUIRE(width % 2 != 0 || ((reinterpret_cast<uintptr_t>(buf) | reinterpret_cast<uintptr_t>(out)) & 15) == 0,
               "soap_interleave_shards: buffers must be 16-byte aligned");
  interleave_shards_kernel<<<num_sms() * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(buf, out, (long long)n_shards,
                                                                                        (long long)rows, (long long)width);
  return check_launch("interleave_shards_kernel");
}

extern "C" int soap_abi_version(void) { return SOAP_B200_ABI_VERSION; }
extern "C" const char *soap_last_error(void) { return last_error().c_str(); }
extern "C" int64_t soap_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int soap_reserve_sms(int n) {
  SOAP_REQUIRE(n >= 0, "soap_reserve_sms: negative count");
  g_reserved_sms.store(n, std::memory_order_relaxed);
  return SOAP_OK;
}

extern "C" int soap_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (auto *r : g_prof) {
    cudaEventDestroy(r->start);
    cudaEventDestroy(r->stop);
    delete r;
  }
  g_prof.clear();
  g_prof_on.store(on ? 1 : 0);
  return SOAP_OK;
}

extern "C" int soap_profile_read(int kernel_id, float *ms_out_host, int capacity) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  int n = 0;
  for (auto *r : g_prof) {
    if (r->id != kernel_id) continue;
    if (cudaEventSynchronize(r->stop) != cudaSuccess) return set_error(SOAP_ECUDA, "soap_profile_read: event sync failed");
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r->start, r->stop) != cudaSuccess)
      return set_error(SOAP_ECUDA, "soap_profile_read: elapsed time failed");
    if (ms_out_host && n < capacity) ms_out_host[n] = ms;
    ++n;
  }
  return n;
}

extern "C" int soap_fill_uniform(void *out, int64_t n, uint64_t seed, int dtype, void *stream) {
  SOAP_REQUIRE(n >= 0, "soap_fill_uniform: negative size");
  if (n == 0) return SOAP_OK;
  SOAP_REQUIRE(out != nullptr, "soap_fill_uniform: null buffer");
  auto s = static_cast<cudaStream_t>(stream);
  const int grid = num_sms() * 8;
  if (dtype == SOAP_F64)
    fill_uniform_kernel<double><<<grid, 256, 0, s>>>(static_cast<double *>(out), (long long)n, seed);
  else if (dtype == SOAP_F32)
    fill_uniform_kernel<float><<<grid, 256, 0, s>>>(static_cast<float *>(out), (long long)n, seed);
  else
    return set_error(SOAP_EUNSUPPORTED, "soap_fill_uniform: dtype %d", dtype);
  return check_launch("fill_uniform_kernel");
}

// strided host <-> device row copy on a stream (cudaMemcpy2DAsync): the atom tiles of a pinned host
// trajectory [F][A][D] are F rows of (tile atoms * D) elements with a pitch of A * D elements.
// to_device == 2: both ends on devices (a peer-mapped buffer included): the copy engines place a rank's result
// slab [rows][A/G] straight into the [rows][A] array of another GPU, no SM involved
extern "C" int soap_copy_rows(void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t row_bytes,
                              size_t rows, int to_device, void *stream) {
  if (row_bytes == 0 || rows == 0) return SOAP_OK;
  SOAP_REQUIRE(dst && src, "soap_copy_rows: null buffer");
  SOAP_REQUIRE(dst_pitch >= row_bytes && src_pitch >= row_bytes, "soap_copy_rows: pitch smaller than a row");
  SOAP_CUDA(cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, row_bytes, rows,
                              to_device == 2 ? cudaMemcpyDefault : (to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost),
                              static_cast<cudaStream_t>(stream)));
  return SOAP_OK;
}
