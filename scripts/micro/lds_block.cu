// Microbenchmark 2: the timelag consumer's per-block structure without any copies -- which part of the
// block protocol costs shared-memory pipe time?   flags: 1 = remainder batches, 2 = mbarrier wait per
// block (always complete), 4 = partial sums to smem + arrive, 8 = slot/address recomputation per block
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

template <int U>
__device__ __forceinline__ void accum(uint32_t m_addr, uint32_t x_addr, const uint32_t (&y_addr)[4], uint32_t off,
                                      uint32_t ustep, double (&acc)[10]) {
  double2 m[U], x[U], y[4][U];
#pragma unroll
  for (int j = 0; j < U; ++j) {
    m[j] = lds_d2(m_addr + off + j * ustep);
    x[j] = lds_d2(x_addr + off + j * ustep);
#pragma unroll
    for (int k = 0; k < 4; ++k) y[k][j] = lds_d2(y_addr[k] + off + j * ustep);
  }
#pragma unroll
  for (int j = 0; j < U; ++j) {
    const double xm0 = x[j].x * m[j].x, xm1 = x[j].y * m[j].y;
    acc[0] = fma(xm0, x[j].x, acc[0]);
    acc[1] = fma(xm1, x[j].y, acc[1]);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      acc[2 + 2 * k] = fma(xm0, y[k][j].x, acc[2 + 2 * k]);
      acc[3 + 2 * k] = fma(xm1, y[k][j].y, acc[3 + 2 * k]);
    }
  }
}

template <int UB>
__global__ void k(double *out, int nblk, int upp, int stride, int flags, long long *clk) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x >> 5) - ((flags & 16) ? 2 : 0);
  const int R = 128;
  double *red = reinterpret_cast<double *>(sm + (size_t)R * stride);
  uint64_t *bar = reinterpret_cast<uint64_t *>(red + 16 * 5 * 32);
  for (int i = threadIdx.x; i < (R * stride) / 8; i += blockDim.x) reinterpret_cast<double *>(sm)[i] = 1.0 + i * 1e-9;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1 << 19); mbar_init(&bar[2], nw); }
  __syncthreads();
  if (threadIdx.x == 0) mbar_arrive(&bar[0]);  // phase 0 complete forever after
  __syncthreads();
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
  const uint32_t m_addr = base, ustep = 16u * nw, off0 = 16u * warp;
  const int my_units = upp > warp ? (upp - warp + nw - 1) / nw : 0;
  double total = 0;
  if (warp >= nw) {  // spinning warps (the producer / reducer of the real kernel while they wait)
    mbar_wait(&bar[2], 0);
    return;
  }
  long long t0 = clock64();
  for (int b = 0; b < nblk; ++b) {
    const int s = (flags & 8) ? (b % 4) : 2;
    const int slot = s * 32 + lane;
    const uint32_t x_addr = base + (uint32_t)slot * stride;
    uint32_t y_addr[4];
    const int wins[4] = {1, 5, 10, 50};
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      int ys = slot - wins[kk];
      if (ys < 0) ys += R;
      y_addr[kk] = base + (uint32_t)ys * stride;
    }
    double acc[10];
#pragma unroll
    for (int q = 0; q < 10; ++q) acc[q] = 0;
    if (flags & 2) mbar_wait(&bar[0], 0);
    int n = 0;
    for (; n + UB <= my_units; n += UB) accum<UB>(m_addr, x_addr, y_addr, off0 + n * ustep, ustep, acc);
    if (flags & 1) {
      if (UB > 4 && n + 4 <= my_units) { accum<4>(m_addr, x_addr, y_addr, off0 + n * ustep, ustep, acc); n += 4; }
      if (UB > 2 && n + 2 <= my_units) { accum<2>(m_addr, x_addr, y_addr, off0 + n * ustep, ustep, acc); n += 2; }
      if (n < my_units) accum<1>(m_addr, x_addr, y_addr, off0 + n * ustep, ustep, acc);
    }
    if (flags & 4) {
      double *r = red + warp * 5 * 32;
#pragma unroll
      for (int q = 0; q < 5; ++q) r[q * 32 + lane] = acc[2 * q] + acc[2 * q + 1];
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar[1]);
    } else {
#pragma unroll
      for (int q = 0; q < 10; ++q) total += acc[q];
    }
  }
  long long t1 = clock64();
  __syncwarp();
  if (lane == 0) mbar_arrive(&bar[2]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = total;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int UB>
void run(int warps, int upp, int flags) {
  const int stride = 16 * (upp | 1), nblk = 400;
  size_t smem = (size_t)128 * stride + 16 * 5 * 32 * 8 + 64;
  cudaFuncSetAttribute(k<UB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  double *out; long long *clk;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&clk, 148 * 8);
  k<UB><<<148, (warps + ((flags & 16) ? 2 : 0)) * 32, smem>>>(out, 4, upp, stride, flags, clk);
  k<UB><<<148, (warps + ((flags & 16) ? 2 : 0)) * 32, smem>>>(out, nblk, upp, stride, flags, clk);
  cudaError_t e = cudaGetLastError();
  long long h[148];
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  int done = 0;  // units actually processed per block
  for (int w = 0; w < warps; ++w) {
    int mu = upp > w ? (upp - w + warps - 1) / warps : 0;
    done += (flags & 1) ? mu : (mu / UB) * UB;
  }
  printf("warps=%2d batch=%d upp=%3d flags=%2d: %.2f clk per warp-unit  [%s]\n", warps, UB, upp, flags,
         (double)h[0] / ((double)nblk * done), cudaGetErrorString(e));
  cudaFree(out); cudaFree(clk);
}

int main() {
  for (int flags : {15, 31}) {
    run<4>(12, 96, flags);
    run<4>(12, 102, flags);
    run<4>(8, 102, flags);
    run<6>(6, 102, flags);
  }
  return 0;
}
