// Microbenchmark 4: what HBM bandwidth does the TMA engine deliver for the box shapes a time-lag ring can
// use?  X[F][A*D] fp64, one frame row = A * 9792 B; a box takes a few hundred bytes of MANY frame rows
// (rows are ~20 MB apart), so DRAM page locality depends on the contiguous bytes per row and box.
//   mode 0: 2-D map, box = 32 frames x 1648 B (6 pieces per atom row), the round-1 kernel
//   mode 1: 3-D map (16 words, lines, frames) with 128B swizzle, box = NF frames x k lines; a piece is c lines
//           and is streamed stage by stage (k lines each) for every block of NB frames (column pipelining)
// Persistent CTAs, one producer lane, `depth` boxes in flight, a consumer warp that only recycles the stages.
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn enc() {
  void *p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  return reinterpret_cast<EncodeTiledFn>(p);
}
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_arrive(uint64_t *b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t n) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(s32(b)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma2(void *dst, const CUtensorMap *m, int c0, int c1, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(s32(dst)),
               "l"(m), "r"(s32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma3(void *dst, const CUtensorMap *m, int c0, int c1, int c2, uint64_t *bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(s32(dst)),
               "l"(m), "r"(s32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

// mode 0: pieces of `pw` words (8 B), box 32 frames;  mode 1: pieces of c lines, stages of k lines, box NF frames
__global__ void __launch_bounds__(64, 1) stream(const __grid_constant__ CUtensorMap tm, int mode, long long F, long long A,
                                                int row_words, int pw, int c, int k, int NF, int NB, int depth, uint32_t box_bytes,
                                                double *sink) {
  extern __shared__ __align__(1024) unsigned char sm[];
  uint64_t *full = reinterpret_cast<uint64_t *>(sm + (size_t)depth * box_bytes);
  uint64_t *empty = full + depth;
  if (threadIdx.x == 0) {
    for (int s = 0; s < depth; ++s) mbar_init(&full[s], 1), mbar_init(&empty[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long nbox = 0;
  const int pieces = mode == 0 ? (row_words + pw - 1) / pw : (row_words / 16 + c - 1) / c;
  const int nblk = mode == 0 ? (int)((F + 31) / 32) : (int)((F + NB - 1) / NB);
  const int stages = mode == 0 ? 1 : (c + k - 1) / k, bpb = mode == 0 ? 1 : NB / NF;
  for (long long a = blockIdx.x; a < A; a += gridDim.x) nbox += (long long)pieces * nblk * stages * bpb;
  if (warp == 0) {
    if (lane == 0) {
      int s = 0, ph = 0;
      long long n = 0;
      for (long long a = blockIdx.x; a < A; a += gridDim.x)
        for (int p = 0; p < pieces; ++p)
          for (int b = 0; b < nblk; ++b)
            for (int st = 0; st < stages; ++st)
              for (int h = 0; h < bpb; ++h, ++n) {
                if (n >= depth) mbar_wait(&empty[s], (uint32_t)(ph ^ 1));
                mbar_expect(&full[s], box_bytes);
                if (mode == 0)
                  tma2(sm + (size_t)s * box_bytes, &tm, (int)(a * row_words + (long long)p * pw), b * 32, &full[s]);
                else
                  tma3(sm + (size_t)s * box_bytes, &tm, 0, (int)(a * (row_words / 16) + p * c + st * k), b * NB + h * NF, &full[s]);
                if (++s == depth) s = 0, ph ^= 1;
              }
    }
  } else {
    int s = 0, ph = 0;
    double acc = 0;
    for (long long n = 0; n < nbox; ++n) {
      mbar_wait(&full[s], (uint32_t)ph);
      acc += reinterpret_cast<double *>(sm + (size_t)s * box_bytes)[lane];
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      if (++s == depth) s = 0, ph ^= 1;
    }
    if (acc == 123.456) sink[0] = acc;
  }
}

int main() {
  const long long F = 1000, A = 1480, AT = 2048;  // tensor holds 2048 atoms, 1480 are streamed (10 per SM)
  const int D = 1224, row_words = D;
  void *X;
  const size_t bytes = (size_t)F * AT * D * 8;
  if (cudaMalloc(&X, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  cudaMemset(X, 0, bytes);
  double *sink;
  cudaMalloc(&sink, 8);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  struct Cfg { int mode, pw, c, k, NF, NB, depth; };
  const Cfg cfgs[] = {
      {0, 206, 0, 0, 32, 32, 2}, {0, 206, 0, 0, 32, 32, 3}, {0, 154, 0, 0, 32, 32, 3},
      {1, 0, 7, 1, 56, 112, 8},  {1, 0, 7, 1, 56, 112, 14}, {1, 0, 9, 3, 56, 112, 4}, {1, 0, 9, 3, 56, 112, 6},
      {1, 0, 7, 7, 56, 112, 2},  {1, 0, 7, 7, 56, 112, 3},  {1, 0, 9, 9, 56, 112, 2}, {1, 0, 5, 1, 112, 224, 6},
      {1, 0, 5, 5, 112, 224, 2}, {1, 0, 13, 13, 32, 32, 3},
  };
  for (const Cfg &cf : cfgs) {
    CUtensorMap tm;
    memset(&tm, 0, sizeof(tm));
    uint32_t box_bytes;
    CUresult r;
    if (cf.mode == 0) {
      cuuint64_t dims[2] = {(cuuint64_t)AT * row_words, (cuuint64_t)F};
      cuuint64_t str[1] = {(cuuint64_t)AT * row_words * 8};
      cuuint32_t box[2] = {(cuuint32_t)cf.pw, 32}, es[2] = {1, 1};
      r = enc()(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, X, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      box_bytes = 32u * cf.pw * 8;
    } else {
      cuuint64_t dims[3] = {16, (cuuint64_t)AT * row_words / 16, (cuuint64_t)F};
      cuuint64_t str[2] = {128, (cuuint64_t)AT * row_words * 8};
      cuuint32_t box[3] = {16, (cuuint32_t)cf.k, (cuuint32_t)cf.NF}, es[3] = {1, 1, 1};
      r = enc()(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, X, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      box_bytes = 128u * cf.k * cf.NF;
    }
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
    const size_t smem = (size_t)cf.depth * box_bytes + 2 * cf.depth * 8 + 64;
    cudaFuncSetAttribute(stream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float best = 1e30f;
    cudaError_t err = cudaSuccess;
    for (int it = 0; it < 3; ++it) {
      cudaEventRecord(e0);
      stream<<<148, 64, smem>>>(tm, cf.mode, F, A, row_words, cf.pw, cf.c, cf.k, cf.NF, cf.NB, cf.depth, box_bytes, sink);
      cudaEventRecord(e1);
      err = cudaDeviceSynchronize();
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    const double useful = (double)F * A * D * 8;
    printf("mode %d pw=%3d c=%2d k=%2d NF=%3d NB=%3d depth=%2d box=%6u B in flight=%6zu B: %7.3f ms  %7.1f GB/s useful (%.3f of 6538.9)  [%s]\n",
           cf.mode, cf.pw, cf.c, cf.k, cf.NF, cf.NB, cf.depth, box_bytes, (size_t)cf.depth * box_bytes, best,
           useful / best / 1e6, useful / best / 1e6 / 6538.9, cudaGetErrorString(err));
  }
  return 0;
}
