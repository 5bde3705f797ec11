// pairs_dmma.cu -- K5 (fp64): the all-pairs SOAP distance matrix as a tiled FP64 tensor-core GEMM
// with the distance as its epilogue.
//
// Reference semantics: getDistanceBetween (src/SOAPify/classify.py:96-117) with one of the
// distances of src/SOAPify/distances.py; out[i][j] = dist(data[i], spectra[j]).
//
// tcgen05 has no f64 kind: the FP64 tensor path of sm_100 is mma.sync (SASS DMMA), here the
// m16n8k8 shape.  C[128x128] tiles, K slabs of 16 doubles, 3-stage cp.async (LDGSTS.128) pipeline,
// 8 warps as 2x4 with 64x32 warp tiles (64 accumulators per thread).  Both operands are row-major
// with the SOAP dimension contiguous ("TN"), which is exactly the fragment order of mma .row.col.
// Shared-memory rows are padded to 24 doubles so that every 128-bit fragment load is conflict
// free; a lane's two k-values of a fragment are taken from ADJACENT doubles (the k order inside
// a slab is a free permutation as long as A and B use the same one), so fragments load as LDS.128.
// Row norms (needed by the cosine-type distances) come from a one-pass kernel into scratch.
#include "common.cuh"

namespace soapb {

constexpr int kBM = 128, kBN = 128, kBK = 16, kStages = 3;
constexpr int kLd = kBK + 8;  // padded row length in doubles (192 bytes)
constexpr int kDThreads = 256;
constexpr int kStageDoubles = (kBM + kBN) * kLd;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void dmma_16x8x8(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
      : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

// squared L2 norm of every row, one warp per row, double accumulation
template <typename T>
__global__ void __launch_bounds__(256) row_sqnorm_kernel(const T *__restrict__ X, double *__restrict__ out,
                                                        long long rows, int D) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const T *x = X + (size_t)row * D;
  double acc = 0.0;
  for (int k = lane; k < D; k += 32) {
    const double v = (double)x[k];
    acc = fma(v, v, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) out[row] = acc;
}

int launch_row_sqnorm(const void *X, double *out, int64_t rows, int D, int dtype, cudaStream_t s) {
  if (rows == 0) return SOAP_OK;
  const unsigned grid = (unsigned)((rows + 7) / 8);
  if (dtype == SOAP_F64)
    row_sqnorm_kernel<double><<<grid, 256, 0, s>>>(static_cast<const double *>(X), out, rows, D);
  else
    row_sqnorm_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float *>(X), out, rows, D);
  return check_launch("row_sqnorm_kernel");
}

// tile order: panels of 8 tile-columns walked row by row, so that the ~148 tiles in flight share
// a few A and B row blocks in L2
__device__ __forceinline__ void tile_coords(long long t, int tiles_m, int tiles_n, int &tm, int &tn) {
  constexpr int PW = 8;
  const long long per_panel = (long long)PW * tiles_m;
  const int panel = (int)(t / per_panel);
  const int first = panel * PW;
  const int width = min(PW, tiles_n - first);
  const long long r = t - (long long)panel * per_panel;
  tm = (int)(r / width);
  tn = first + (int)(r % width);
}

__global__ void __launch_bounds__(kDThreads, 1)
    pairs_dmma_kernel(const double *__restrict__ A, const double *__restrict__ B, double *__restrict__ out,
                      const double *__restrict__ uu, const double *__restrict__ vv, long long M, long long N,
                      int D, long long ldo, int kind, int power, int tiles_m, int tiles_n) {
  extern __shared__ __align__(128) double smem_d[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, tig = lane & 3;
  const int wm = warp >> 2, wn = warp & 3;  // 2 x 4 warps, 64 x 32 each
  int tm, tn;
  tile_coords(blockIdx.x, tiles_m, tiles_n, tm, tn);
  const long long m0 = (long long)tm * kBM, n0 = (long long)tn * kBN;
  const int KT = (D + kBK - 1) / kBK;
  const uint32_t smem_base = smem_u32(smem_d);

  // loader mapping: 8 chunks of 16 bytes per 128-byte row slab; thread -> (row, chunk) x 4 passes
  const int lrow = tid >> 3, lchunk = tid & 7;
  auto load_stage = [&](int kt, int slot) {
    const int k = kt * kBK + lchunk * 2;
    const int nbytes = (k < D) ? 16 : 0;  // D is even: a chunk is either fully valid or zero filled
    const int kc = (k < D) ? k : 0;
    const uint32_t sA = smem_base + (uint32_t)(slot * kStageDoubles) * 8u;
    const uint32_t sB = sA + (uint32_t)(kBM * kLd) * 8u;
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const int r = lrow + pass * 32;
      long long ra = m0 + r, rb = n0 + r;
      ra = ra < M ? ra : M - 1;  // out-of-range rows replay the last row; their results are not stored
      rb = rb < N ? rb : N - 1;
      cp_async16(sA + (uint32_t)(r * kLd + lchunk * 2) * 8u, A + (size_t)ra * D + kc, nbytes);
      cp_async16(sB + (uint32_t)(r * kLd + lchunk * 2) * 8u, B + (size_t)rb * D + kc, nbytes);
    }
  };

  double acc[4][4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[i][j][v] = 0.0;

#pragma unroll
  for (int s = 0; s < kStages - 1; ++s) {
    if (s < KT) load_stage(s, s);
    cp_async_commit();
  }

  for (int kt = 0; kt < KT; ++kt) {
    cp_async_wait<kStages - 2>();
    __syncthreads();  // slab kt has landed for everyone; slot (kt-1)%S is free for everyone
    {
      const int nk = kt + kStages - 1;
      if (nk < KT) load_stage(nk, nk % kStages);
      cp_async_commit();
    }
    const uint32_t sA = smem_base + (uint32_t)((kt % kStages) * kStageDoubles) * 8u;
    const uint32_t sB = sA + (uint32_t)(kBM * kLd) * 8u;
#pragma unroll
    for (int ks = 0; ks < kBK / 8; ++ks) {
      double a[4][4], b[4][2];
      const uint32_t koff = (uint32_t)(ks * 8 + 2 * tig) * 8u;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t row = (uint32_t)(wm * 64 + i * 16 + g);
        // (a0,a2) = row g, (a1,a3) = row g+8: two adjacent doubles play k = tig and k = tig+4
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a[i][0]), "=d"(a[i][2]) : "r"(sA + row * (kLd * 8u) + koff));
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(a[i][1]), "=d"(a[i][3]) : "r"(sA + (row + 8) * (kLd * 8u) + koff));
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t col = (uint32_t)(wn * 32 + j * 8 + g);
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(b[j][0]), "=d"(b[j][1]) : "r"(sB + col * (kLd * 8u) + koff));
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma_16x8x8(acc[i][j], a[i], b[j]);
    }
  }
  cp_async_wait<0>();

  // epilogue: distance from (uv, uu_i, vv_j), streamed out as 16-byte stores
  const bool need_norms = (kind != SOAP_KIND_NORMALIZED);
  const bool vec_ok = ((ldo & 1) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const long long row = m0 + wm * 64 + i * 16 + g + h * 8;
      if (row >= M) continue;
      const double ui = need_norms ? uu[row] : 1.0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const long long col = n0 + wn * 32 + j * 8 + 2 * tig;
        if (col >= N) continue;
        const double d0 = soap_epilogue(acc[i][j][2 * h], ui, need_norms ? vv[col] : 1.0, kind, power);
        double *dst = out + (size_t)row * ldo + col;
        if (col + 1 < N) {
          const double d1 = soap_epilogue(acc[i][j][2 * h + 1], ui, need_norms ? vv[col + 1] : 1.0, kind, power);
          if (vec_ok) {
            asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(dst), "d"(d0), "d"(d1) : "memory");
          } else {
            dst[0] = d0;
            dst[1] = d1;
          }
        } else {
          dst[0] = d0;
        }
      }
    }
  }
}

size_t pairs_dmma_scratch(int64_t M, int64_t N, int D) { return (size_t)(M + N) * sizeof(double) + 256; }

int pairs_dmma(const double *A, const double *B, double *out, int64_t M, int64_t N, int D, int64_t ldo,
               int kind, int power, void *scratch, cudaStream_t s) {
  SOAP_REQUIRE(out != nullptr, "soap_pairs: the tensor-core path writes the full matrix (out == NULL)");
  SOAP_REQUIRE(D % 2 == 0, "soap_pairs(tensor, fp64): the SOAP dimension must be even");
  SOAP_REQUIRE(((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0,
               "soap_pairs(tensor, fp64): inputs must be 16-byte aligned");
  double *uu = static_cast<double *>(scratch);
  double *vv = uu + M;
  if (kind != SOAP_KIND_NORMALIZED) {
    int rc = launch_row_sqnorm(A, uu, M, D, SOAP_F64, s);
    if (rc) return rc;
    if (A == B && M == N) {
      vv = uu;
    } else {
      rc = launch_row_sqnorm(B, vv, N, D, SOAP_F64, s);
      if (rc) return rc;
    }
  }
  const int tiles_m = (int)((M + kBM - 1) / kBM), tiles_n = (int)((N + kBN - 1) / kBN);
  const long long tiles = (long long)tiles_m * tiles_n;
  SOAP_REQUIRE(tiles < 2147483647LL, "soap_pairs: too many tiles");
  const size_t smem = (size_t)kStages * kStageDoubles * sizeof(double);
  SOAP_CUDA(cudaFuncSetAttribute(pairs_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ProfScope prof(SOAP_PROF_PAIRS, s);
  pairs_dmma_kernel<<<(unsigned)tiles, kDThreads, smem, s>>>(A, B, out, uu, vv, M, N, D, ldo, kind, power, tiles_m,
                                                             tiles_n);
  return check_launch("pairs_dmma_kernel");
}

}  // namespace soapb
