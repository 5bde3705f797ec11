// Microbenchmark 3: consumer loop of the register-blocked time-lag sweep (no copies).
// Geometry: a lane owns J consecutive frames of one 16-byte unit (chunk) of a 128-byte line; a half warp
// = 16 such groups (B = 16*J frames per block), the two halves of a warp take neighbouring chunks, CW
// warps cover the 8 chunks of a line.  The ring is [line][R frame slots][128 B] with the TMA 128B swizzle
// (chunk ^= slot & 7), so the 8 lanes of a quarter warp (8 consecutive groups, J odd) hit 8 distinct
// 16-byte bank groups.  Per unit a lane loads J (x) + 10 (frames f-1..f-10) + J (lag 50) words instead of
// 5 per frame.  flags: 1 = partial sums to smem per block, 2 = mbarrier wait per line stage
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ double2 lds_d2(uint32_t a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void mbar_init(uint64_t *b, int n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(n));
}
__device__ __forceinline__ void mbar_arrive(uint64_t *b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"((uint32_t)__cvta_generic_to_shared(b)), "r"(parity) : "memory");
  } while (!ok);
}

constexpr int J = 7, G = 16, B = G * J, R = 168, Q = 5;

template <int CW>
__global__ void __launch_bounds__(CW * 32, 1) k(double *out, int nblk, int c, int flags, long long *clk) {
  extern __shared__ __align__(1024) unsigned char sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int half = lane >> 4, g = lane & 15;
  const int chunk = (2 * warp + half) & 7;  // CW = 8: warps 4..7 take the odd lines
  const int lsel = (2 * warp + half) >> 3, lstep = CW / 4;
  unsigned char *msm = sm + (size_t)c * R * 128;
  double *red = reinterpret_cast<double *>(msm + (size_t)c * 128);  // [CW][2][Q][B]
  uint64_t *bar = reinterpret_cast<uint64_t *>(red + (size_t)CW * 2 * Q * B);
  for (int i = threadIdx.x; i < (c * R * 128 + c * 128) / 8; i += blockDim.x) reinterpret_cast<double *>(sm)[i] = 1.0 + i * 1e-9;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1 << 19); }
  __syncthreads();
  if (threadIdx.x == 0) mbar_arrive(&bar[0]);
  __syncthreads();
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm), mbase = (uint32_t)__cvta_generic_to_shared(msm);
  double total = 0;
  long long t0 = clock64();
  for (int b = 0; b < nblk; ++b) {
    const int s0 = (b * B) % R + g * J;
    uint32_t ax[J], ah[10], al[J];
    auto addr = [&](int slot) {
      slot = slot < 0 ? slot + R : (slot >= R ? slot - R : slot);
      return (uint32_t)(slot * 128 + ((chunk ^ (slot & 7)) << 4));
    };
#pragma unroll
    for (int i = 0; i < J; ++i) ax[i] = addr(s0 + i), al[i] = addr(s0 + i - 50);
#pragma unroll
    for (int i = 0; i < 10; ++i) ah[i] = addr(s0 - 1 - i);
    double acc[Q][J];
#pragma unroll
    for (int q = 0; q < Q; ++q)
#pragma unroll
      for (int i = 0; i < J; ++i) acc[q][i] = 0;
    for (int l = lsel; l < c; l += lstep) {
      if (flags & 2) mbar_wait(&bar[0], 0);
      const uint32_t rb = base + (uint32_t)l * R * 128;
      double2 x[J], h[10], y[J];
      const double2 m = lds_d2(mbase + (l * 8 + chunk) * 16);
#pragma unroll
      for (int i = 0; i < J; ++i) x[i] = lds_d2(rb + ax[i]);
#pragma unroll
      for (int i = 0; i < 10; ++i) h[i] = lds_d2(rb + ah[i]);
#pragma unroll
      for (int i = 0; i < J; ++i) y[i] = lds_d2(rb + al[i]);
#pragma unroll
      for (int i = 0; i < J; ++i) {
        const double xm0 = x[i].x * m.x, xm1 = x[i].y * m.y;
        const double2 p1 = i >= 1 ? x[i - 1] : h[0 - i];
        const double2 p5 = i >= 5 ? x[i - 5] : h[4 - i];
        const double2 p10 = h[9 - i];
        acc[0][i] = fma(xm1, x[i].y, fma(xm0, x[i].x, acc[0][i]));
        acc[1][i] = fma(xm1, p1.y, fma(xm0, p1.x, acc[1][i]));
        acc[2][i] = fma(xm1, p5.y, fma(xm0, p5.x, acc[2][i]));
        acc[3][i] = fma(xm1, p10.y, fma(xm0, p10.x, acc[3][i]));
        acc[4][i] = fma(xm1, y[i].y, fma(xm0, y[i].x, acc[4][i]));
      }
    }
    if (flags & 1) {
      double *r = red + (size_t)(warp * 2 + half) * Q * B;
#pragma unroll
      for (int q = 0; q < Q; ++q)
#pragma unroll
        for (int i = 0; i < J; ++i) r[q * B + g * J + i] = acc[q][i];
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar[1]);
    } else {
#pragma unroll
      for (int q = 0; q < Q; ++q)
#pragma unroll
        for (int i = 0; i < J; ++i) total += acc[q][i];
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = total;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int CW>
void run(int c, int flags) {
  const int nblk = 300;
  size_t smem = (size_t)c * R * 128 + c * 128 + (size_t)CW * 2 * Q * B * 8 + 64;
  cudaFuncSetAttribute(k<CW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  double *out; long long *clk;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&clk, 148 * 8);
  k<CW><<<148, CW * 32, smem>>>(out, 4, c, flags, clk);
  k<CW><<<148, CW * 32, smem>>>(out, nblk, c, flags, clk);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  // one line stage = 8 chunks x B frames = B/4 units of 32 frames
  const double stages = (double)nblk * c;
  printf("J=%d warps=%d lines=%d flags=%d smem=%zu: %.1f clk per line stage = %.2f clk per (unit x 32 frames)  [%s]\n", J, CW,
         c, flags, smem, (double)h[0] / stages, (double)h[0] / stages / (B / 4.0), cudaGetErrorString(e));
  cudaFree(out); cudaFree(clk);
}

int main() {
  for (int flags : {0, 1, 3}) {
    run<4>(7, flags);
    run<4>(8, flags);
    run<8>(8, flags);
  }
  return 0;
}
