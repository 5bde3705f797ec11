// Microbenchmark / probe: which start coordinates does a tiled TMA box accept?  (rows that are not
// 16-byte multiples.)  A 2-D tensor [rows][cols] is described once with 4-byte and once with 8-byte elements; boxes
// are fetched at every start column 0..7 and compared with the source.  A faulting coordinate shows up as a CUDA error.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_align tma_align.cu -lcuda
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <typename T>
__global__ void k(const __grid_constant__ CUtensorMap map, int c0, int box_cols, int box_rows, T *out) {
  extern __shared__ __align__(128) unsigned char sm[];
  uint64_t *bar = reinterpret_cast<uint64_t *>(sm + 32768);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t bytes = (uint32_t)(box_cols * box_rows * sizeof(T));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(sm)),
                 "l"(&map), "r"(smem_u32(bar)), "r"(c0), "r"(0)
                 : "memory");
  }
  uint32_t ok = 0;
  for (int i = 0; i < (1 << 22) && !ok; ++i)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
  if (!ok) __trap();
  for (int i = threadIdx.x; i < box_cols * box_rows; i += blockDim.x) out[i] = reinterpret_cast<T *>(sm)[i];
}

typedef CUresult (*Enc)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                        const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <typename T>
void probe(Enc enc, CUtensorMapDataType dt, const char *name, int cols) {
  const int rows = 16, box_cols = 32 / (int)sizeof(T) * 4, box_rows = 8;  // box rows of 128 bytes
  std::vector<T> h((size_t)rows * cols);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (T)(i + 1);
  T *d, *o;
  cudaMalloc(&d, h.size() * sizeof(T));
  cudaMalloc(&o, box_cols * box_rows * sizeof(T));
  cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
  CUtensorMap map;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)cols * sizeof(T)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows}, es[2] = {1, 1};
  CUresult r = enc(&map, dt, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                   CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("%s cols=%d (row pitch %zu B): encode rc=%d\n", name, cols, cols * sizeof(T), (int)r);
  if (r != CUDA_SUCCESS) return;
  cudaFuncSetAttribute(k<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  for (int c0 = 0; c0 < 8; ++c0) {
    cudaMemset(o, 0, box_cols * box_rows * sizeof(T));
    k<T><<<1, 128, 40000>>>(map, c0, box_cols, box_rows, o);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("  start col %d (byte %zu): CUDA error %s  -> aborting this probe\n", c0, c0 * sizeof(T), cudaGetErrorString(e));
      return;  // (a fault poisons the context)
    }
    std::vector<T> g(box_cols * box_rows);
    cudaMemcpy(g.data(), o, g.size() * sizeof(T), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int rr = 0; rr < box_rows; ++rr)
      for (int cc = 0; cc < box_cols; ++cc) {
        const T want = (c0 + cc < cols) ? h[(size_t)rr * cols + c0 + cc] : (T)0;
        bad += g[rr * box_cols + cc] != want;
      }
    printf("  start col %d (byte %zu): %s\n", c0, c0 * sizeof(T), bad ? "WRONG DATA" : "ok");
  }
  cudaFree(d);
  cudaFree(o);
}

int main(int argc, char **argv) {
  void *p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return 1;
  Enc enc = (Enc)p;
  const int which = argc > 1 ? atoi(argv[1]) : 0;
  if (which == 0) probe<float>(enc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, "float32 elements", 1000);
  if (which == 1) probe<unsigned long long>(enc, CU_TENSOR_MAP_DATA_TYPE_UINT64, "uint64 elements", 500);
  if (which == 2) probe<float>(enc, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, "float32 elements, pitch 4 mod 16", 1001);
  return 0;
}
