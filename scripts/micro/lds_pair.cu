// Microbenchmark 3: the consumer loop of a "two frames per lane" sweep (lane = frame pair 2p, 2p+1; even and odd
// frames in separate ring halves so that consecutive lanes read consecutive slots; the lag-1 partner of the odd
// frame is the lane's own even frame) against the lane = frame loop of lds_block.cu, no copies.  Windows 1,5,10,50.
// Reports clk per (16-byte unit x 32 frames), the unit of DESIGN 4.2.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o lds_pair lds_pair.cu
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

// U units of a frame pair: 9 loads per unit (x even, x odd, lag-1 partner of the even frame, 2 x 3 partners)
template <int U>
__device__ __forceinline__ void accum_pair(uint32_t m_addr, uint32_t xe, uint32_t xo, uint32_t y1, const uint32_t (&ye)[3],
                                           const uint32_t (&yo)[3], uint32_t off, uint32_t ustep, double (&acc)[20]) {
  double2 m[U], a[U], b[U], p1[U], pe[3][U], po[3][U];
#pragma unroll
  for (int j = 0; j < U; ++j) {
    const uint32_t o = off + j * ustep;
    m[j] = lds_d2(m_addr + o);
    a[j] = lds_d2(xe + o);
    b[j] = lds_d2(xo + o);
    p1[j] = lds_d2(y1 + o);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      pe[k][j] = lds_d2(ye[k] + o);
      po[k][j] = lds_d2(yo[k] + o);
    }
  }
#pragma unroll
  for (int j = 0; j < U; ++j) {
    const double am0 = a[j].x * m[j].x, am1 = a[j].y * m[j].y, bm0 = b[j].x * m[j].x, bm1 = b[j].y * m[j].y;
    acc[0] = fma(am0, a[j].x, acc[0]);
    acc[1] = fma(am1, a[j].y, acc[1]);
    acc[10] = fma(bm0, b[j].x, acc[10]);
    acc[11] = fma(bm1, b[j].y, acc[11]);
    acc[2] = fma(am0, p1[j].x, acc[2]);  // even frame, lag 1
    acc[3] = fma(am1, p1[j].y, acc[3]);
    acc[12] = fma(bm0, a[j].x, acc[12]);  // odd frame, lag 1: the partner is the lane's own even frame
    acc[13] = fma(bm1, a[j].y, acc[13]);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      acc[4 + 2 * k] = fma(am0, pe[k][j].x, acc[4 + 2 * k]);
      acc[5 + 2 * k] = fma(am1, pe[k][j].y, acc[5 + 2 * k]);
      acc[14 + 2 * k] = fma(bm0, po[k][j].x, acc[14 + 2 * k]);
      acc[15 + 2 * k] = fma(bm1, po[k][j].y, acc[15 + 2 * k]);
    }
  }
}

template <int UB>
__global__ void k(double *out, int nblk, int upp, int stride, int RP, long long *clk) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  // two half rings of RP pair slots each
  double *red = reinterpret_cast<double *>(sm + (size_t)2 * RP * stride);
  uint64_t *bar = reinterpret_cast<uint64_t *>(red + 16 * 10 * 32);
  for (int i = threadIdx.x; i < (2 * RP * stride) / 8; i += blockDim.x) reinterpret_cast<double *>(sm)[i] = 1.0 + i * 1e-9;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1 << 19); }
  __syncthreads();
  if (threadIdx.x == 0) mbar_arrive(&bar[0]);
  __syncthreads();
  const uint32_t base_e = (uint32_t)__cvta_generic_to_shared(sm), base_o = base_e + (uint32_t)RP * stride;
  const uint32_t m_addr = base_e, ustep = 16u * nw, off0 = 16u * warp;
  const int my_units = upp > warp ? (upp - warp + nw - 1) / nw : 0;
  double total = 0;
  long long t0 = clock64();
  for (int b = 0; b < nblk; ++b) {
    const int p = (b % 3) * 32 + lane;  // pair slot of this lane (RP = 96)
    auto slot = [&](int q) { q %= RP; return q < 0 ? q + RP : q; };
    const uint32_t xe = base_e + (uint32_t)slot(p) * stride, xo = base_o + (uint32_t)slot(p) * stride;
    const uint32_t y1 = base_o + (uint32_t)slot(p - 1) * stride;
    // lag 5: 2p-5 odd (p-3), 2p-4 even (p-2); lag 10: 2p-10 even (p-5), 2p-9 odd (p-5); lag 50: p-25 / p-25
    const uint32_t ye[3] = {base_o + (uint32_t)slot(p - 3) * stride, base_e + (uint32_t)slot(p - 5) * stride, base_e + (uint32_t)slot(p - 25) * stride};
    const uint32_t yo[3] = {base_e + (uint32_t)slot(p - 2) * stride, base_o + (uint32_t)slot(p - 5) * stride, base_o + (uint32_t)slot(p - 25) * stride};
    double acc[20];
#pragma unroll
    for (int q = 0; q < 20; ++q) acc[q] = 0;
    mbar_wait(&bar[0], 0);
    int n = 0;
    for (; n + UB <= my_units; n += UB) accum_pair<UB>(m_addr, xe, xo, y1, ye, yo, off0 + n * ustep, ustep, acc);
    for (; n < my_units; ++n) accum_pair<1>(m_addr, xe, xo, y1, ye, yo, off0 + n * ustep, ustep, acc);
    double *r = red + warp * 10 * 32;
#pragma unroll
    for (int q = 0; q < 10; ++q) r[q * 32 + lane] = acc[2 * q] + acc[2 * q + 1];
    __syncwarp();
    if (lane == 0) mbar_arrive(&bar[1]);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = total;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <int UB>
void run(int warps, int upp) {
  const int stride = 16 * (upp | 1), nblk = 300, RP = 96;
  size_t smem = (size_t)2 * RP * stride + 16 * 10 * 32 * 8 + 64;
  cudaFuncSetAttribute(k<UB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  double *out; long long *clk;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&clk, 148 * 8);
  k<UB><<<148, warps * 32, smem>>>(out, 4, upp, stride, RP, clk);
  k<UB><<<148, warps * 32, smem>>>(out, nblk, upp, stride, RP, clk);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  // a block = 64 frames x upp units = 2 * upp (unit x 32 frames)
  printf("pairs: warps=%2d batch=%d upp=%3d smem=%zu: %.2f clk per (unit x 32 frames)  [%s]\n", warps, UB, upp, smem,
         (double)h[0] / ((double)nblk * 2 * upp), cudaGetErrorString(e));
  cudaFree(out); cudaFree(clk);
}

int main() {
  run<2>(8, 61);  // (2 x 96 pair slots x 61 units = 187 kB: the widest piece a double-buffered 64-frame block leaves room for)
  run<1>(8, 61);
  run<3>(8, 61);
  run<2>(6, 61);
  run<2>(10, 61);
  run<2>(8, 55);
  run<4>(6, 61);  // 8 warps per CTA in all: 255 registers per thread, room for deeper batches
  run<6>(6, 61);
  run<3>(6, 61);
  run<4>(4, 61);
  return 0;
}
