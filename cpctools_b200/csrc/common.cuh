// common.cuh -- shared device/host helpers for libsoapb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <string>

#include "../../include/soap_b200.h"

namespace soapb {

// ---------------------------------------------------------------------------------------
// host-side error plumbing (no exception crosses the C ABI)
// ---------------------------------------------------------------------------------------
std::string &last_error();
extern std::atomic<int64_t> g_launches;
int set_error(int code, const char *fmt, ...);
int check_launch(const char *what);
int num_sms();
int persistent_sms();  // num_sms() minus soap_reserve_sms()
// profiling hook: RAII bracket of one kernel launch with CUDA events (no-op unless enabled)
struct ProfScope {
  int id;
  cudaStream_t stream;
  void *rec;
  ProfScope(int kernel_id, cudaStream_t s);
  ~ProfScope();
};
int max_smem_optin();

#define SOAP_REQUIRE(cond, ...)                                  \
  do {                                                           \
    if (!(cond)) return soapb::set_error(SOAP_EINVAL, __VA_ARGS__); \
  } while (0)

#define SOAP_CUDA(call)                                                                    \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess)                                                                \
      return soapb::set_error(SOAP_ECUDA, "%s failed: %s", #call, cudaGetErrorString(e__)); \
  } while (0)

// ---------------------------------------------------------------------------------------
// device helpers: mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP)
// ---------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  // make barrier initialisation visible to the async (TMA) proxy
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  // order generic-proxy smem accesses before subsequent async-proxy (TMA) writes
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Wait with a watchdog: a hand-shake that cannot complete (a protocol bug, a faulted bulk copy) traps
// after ~20 s instead of hanging the GPU.  The common case returns from the first poll.
static __device__ __noinline__ void mbar_wait_slow(uint64_t *bar, uint32_t parity) {
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    for (int i = 0; i < 4096; ++i)
      if (mbar_try_wait(bar, parity)) return;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 20000000000ull) __trap();
  }
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  mbar_wait_slow(bar, parity);
}

// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes,
                                         uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// streaming (read-once) 128-bit global accesses
__device__ __forceinline__ int4 ldg_stream16(const void *p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream16(void *p, const int4 &v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

// ---- shared-memory access by 32-bit address (keeps the hot loop free of 64-bit pointer math) ----
__device__ __forceinline__ double2 lds_d2(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// distance epilogue shared by the time-lag and the pair kernels.  Mirrors, in double:
//   SIMPLE      sqrt(2 - 2 (1 - clip(1 - uv / sqrt(uu vv), 0, 2)))   distances.py:18,33 + scipy
//   KERNEL      sqrt(2 - 2 (uv / (sqrt(uu) sqrt(vv)))^n)             distances.py:51,66
//   NORMALIZED  sqrt(|2 - 2 uv|)                                     distances.py:83
//   DOT         the kernel value itself;  COSCLIP: simpleKernelSoap;  EUCLID: sqrt(sum (x-y)^2)
__device__ __forceinline__ double ipow(double b, int n) {
  // b**n for the small positive integer powers SOAPdistance is used with; n <= 0 -> pow()
  if (n == 1) return b;
  if (n <= 0) return pow(b, (double)n);
  double r = 1.0;
  double s = b;
  int e = n;
  while (e) {
    if (e & 1) r *= s;
    s *= s;
    e >>= 1;
  }
  return r;
}
__device__ __forceinline__ double soap_epilogue(double uv, double uu, double vv, int kind,
                                                int power) {
  switch (kind) {
    case SOAP_KIND_SIMPLE: {
      double cosd = 1.0 - uv / sqrt(uu * vv);
      cosd = (cosd != cosd) ? cosd : fmin(fmax(cosd, 0.0), 2.0);  // numpy.clip keeps NaN
      return sqrt(2.0 - 2.0 * (1.0 - cosd));
    }
    case SOAP_KIND_KERNEL:
      return sqrt(2.0 - 2.0 * ipow(uv / (sqrt(uu) * sqrt(vv)), power));
    case SOAP_KIND_NORMALIZED:
      return sqrt(fabs(2.0 - 2.0 * uv));
    case SOAP_KIND_NORMALIZED_FUSED: {
      const double nu = (uu == 0.0) ? 1.0 : sqrt(uu), nv = (vv == 0.0) ? 1.0 : sqrt(vv);
      return sqrt(fabs(2.0 - 2.0 * (uv / (nu * nv))));
    }
    case SOAP_KIND_COSCLIP: {
      double cosd = 1.0 - uv / sqrt(uu * vv);
      cosd = (cosd != cosd) ? cosd : fmin(fmax(cosd, 0.0), 2.0);
      return 1.0 - cosd;
    }
    case SOAP_KIND_EUCLID:  // uv carries sum (x - y)^2
      return sqrt(uv);
    default:  // SOAP_KIND_DOT
      return ipow(uv / (sqrt(uu) * sqrt(vv)), power);
  }
}
#endif  // __CUDACC__

}  // namespace soapb
