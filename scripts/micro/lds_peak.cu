// Microbenchmark: shared-memory read throughput of the timelag consumer's access pattern
// (lane = ring slot, slots 16*odd bytes apart, 1 + NW 128-bit loads per unit + one broadcast load).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o lds_peak lds_peak.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ double2 lds_d2(uint32_t a) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ double lds_d1(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
  return v;
}

template <int MODE>  // 0: 6x LDS.128 per unit (x, 4 y, m)   1: same bytes with LDS.64   2: loads only, no FMA
__global__ void k(double *out, int iters, int upp, int stride, long long *clk) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int i = threadIdx.x; i < (128 * stride) / 8; i += blockDim.x) reinterpret_cast<double *>(sm)[i] = 1.0 + i * 1e-9;
  __syncthreads();
  uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
  const uint32_t m_addr = base;  // broadcast region = slot 0
  uint32_t x_addr = base + (uint32_t)(64 + lane) * stride;
  uint32_t y_addr[4] = {x_addr - 1u * stride, x_addr - 5u * stride, x_addr - 10u * stride, x_addr - 50u * stride};
  double acc[10];
  for (int q = 0; q < 10; ++q) acc[q] = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    for (int u0 = warp; u0 + 3 * nw < upp; u0 += 4 * nw) {
      if (MODE == 1) {
        double m[8], x[8], y[4][8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t off = 16u * (u0 + j * nw);
          m[2 * j] = lds_d1(m_addr + off); m[2 * j + 1] = lds_d1(m_addr + off + 8);
          x[2 * j] = lds_d1(x_addr + off); x[2 * j + 1] = lds_d1(x_addr + off + 8);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) { y[kk][2 * j] = lds_d1(y_addr[kk] + off); y[kk][2 * j + 1] = lds_d1(y_addr[kk] + off + 8); }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const double xm = x[j] * m[j];
          acc[j & 1] = fma(xm, x[j], acc[j & 1]);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) acc[2 + 2 * kk + (j & 1)] = fma(xm, y[kk][j], acc[2 + 2 * kk + (j & 1)]);
        }
      } else {
        double2 m[4], x[4], y[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint32_t off = 16u * (u0 + j * nw);
          m[j] = lds_d2(m_addr + off);
          x[j] = lds_d2(x_addr + off);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) y[kk][j] = lds_d2(y_addr[kk] + off);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (MODE == 2) {
            acc[0] += m[j].x + x[j].y;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) acc[1 + kk] += y[kk][j].x;
          } else {
            const double xm0 = x[j].x * m[j].x, xm1 = x[j].y * m[j].y;
            acc[0] = fma(xm0, x[j].x, acc[0]);
            acc[1] = fma(xm1, x[j].y, acc[1]);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              acc[2 + 2 * kk] = fma(xm0, y[kk][j].x, acc[2 + 2 * kk]);
              acc[3 + 2 * kk] = fma(xm1, y[kk][j].y, acc[3 + 2 * kk]);
            }
          }
        }
      }
    }
  }
  long long t1 = clock64();
  double s = 0;
  for (int q = 0; q < 10; ++q) s += acc[q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int warps, int upp) {
  const int stride = 16 * (upp | 1), iters = 200;
  size_t smem = (size_t)128 * stride;
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  double *out; long long *clk;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&clk, 148 * 8);
  k<MODE><<<148, warps * 32, smem>>>(out, 2, upp, stride, clk);
  k<MODE><<<148, warps * 32, smem>>>(out, iters, upp, stride, clk);
  long long h[148];
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError(); cudaDeviceSynchronize();
  const int per_warp_groups = (upp / (4 * warps));  // groups of 4 units per warp per iteration
  const double units = (double)iters * per_warp_groups * 4 * warps;  // warp-units per CTA
  printf("%-18s warps=%2d upp=%3d: %.2f clk per warp-unit (ideal 21 wavefronts: 4x5 + 1)  [%s]\n", name, warps, upp,
         (double)h[0] / units, cudaGetErrorString(e));
  cudaFree(out); cudaFree(clk);
}

int main() {
  for (int w : {4, 8, 12, 16}) {
    run<0>("LDS.128 + DFMA", w, 96);
    run<1>("LDS.64 + DFMA", w, 96);
    run<2>("LDS.128 only", w, 96);
  }
  return 0;
}
