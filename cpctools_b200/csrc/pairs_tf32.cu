// pairs_tf32.cu -- K5 (fp32): the all-pairs SOAP distance matrix as a tcgen05 tensor-core GEMM with
// 3xTF32 error compensation and the distance as its epilogue.
//
// Reference semantics: getDistanceBetween (src/SOAPify/classify.py:96-117) on float32 data with one
// of the distances of src/SOAPify/distances.py; output float32 (classify.py:113).
//
// G = A B^T is computed as  Ahi Bhi^T + Ahi Blo^T + Alo Bhi^T  with hi = tf32(x) (round to nearest)
// and lo = tf32(x - hi): three tcgen05.mma.kind::tf32 passes accumulate into ONE fp32 TMEM
// accumulator.  Per product this restores ~2^-21 relative accuracy, BUT the accumulator adds every
// K = 8 slice with fp32 round-toward-zero: over 72 K-slices x 3 passes the one-sided bias reaches
// 6e-6 on the kernel value (measured: max |d^2 - d^2_ref| = 1.2e-5), which is OUTSIDE the 1e-6
// tolerance of the fp32 path.  This kernel is therefore the opt-in SOAP_PAIRS_TENSOR_NATIVE path
// (held to 5e-5 in the tests); the default fp32 tensor path is the exact INT8 scheme of
// pairs_i8.cu, which is also faster.
//
// One persistent CTA per SM, warp specialised:
//   warp 0      TMA producer: 4 tensor-map loads per K block (Ahi, Alo, Bhi, Blo; 128-byte swizzle)
//   warp 1      MMA issuer (one elected lane): 4 K-steps x 3 passes of M128 x N256 x K8, operands
//               straight from shared memory (UMMA descriptors), accumulator in TMEM;
//               tcgen05.commit releases the smem stage / publishes the accumulator
//   warps 2..5  epilogue: tcgen05.ld 32 lanes x 32 columns, distance from (uv, uu_i, vv_j),
//               128-byte row segments streamed to global memory
// The 512 TMEM columns hold two 256-column accumulators, so the epilogue of tile t overlaps the
// mainloop of tile t+1.  hi/lo splits and row norms are produced once by a streaming pre-pass.
#include "tc05.cuh"

namespace soapb {

int launch_row_sqnorm(const void *X, double *out, int64_t rows, int D, int dtype, cudaStream_t s);

namespace tf {
constexpr int BM = 128, BN = 256, BK = 32, UK = 8, STAGES = 2;
constexpr int A_TILE = BM * BK * 4;                       // 16 KB
constexpr int B_TILE = BN * BK * 4;                       // 32 KB
constexpr int STAGE_BYTES = 2 * A_TILE + 2 * B_TILE;      // hi + lo of both operands: 96 KB
constexpr int THREADS = 192;
constexpr int TMEM_COLS = 512;
constexpr size_t SMEM_BYTES = 1024 + (size_t)STAGES * STAGE_BYTES + 256;

// instruction descriptor: D = f32, A = B = tf32, both K-major, N = 256, M = 128
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}  // namespace tf

// ---- pre-pass: hi/lo TF32 split (zero padded to Dp) + squared row norms -------------------------------
__global__ void __launch_bounds__(256)
    split_tf32_kernel(const float *__restrict__ X, float *__restrict__ hi, float *__restrict__ lo,
                      double *__restrict__ norm2, long long rows, int D, int Dp) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float *x = X + (size_t)row * D;
  float *h = hi + (size_t)row * Dp, *l = lo + (size_t)row * Dp;
  double acc = 0.0;
  for (int k = lane; k < Dp; k += 32) {
    const float v = (k < D) ? x[k] : 0.f;
    uint32_t hb, lb;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(v));
    const float r = v - __uint_as_float(hb);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"(r));
    h[k] = __uint_as_float(hb);
    l[k] = __uint_as_float(lb);
    acc = fma((double)v, (double)v, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) norm2[row] = acc;
}

__device__ __forceinline__ void tile_coords_tf(long long t, int tiles_m, int tiles_n, int &tm, int &tn) {
  constexpr int PW = 4;  // panels of 4 x 256 columns
  const long long per_panel = (long long)PW * tiles_m;
  const int panel = (int)(t / per_panel);
  const int first = panel * PW;
  const int width = min(PW, tiles_n - first);
  const long long r = t - (long long)panel * per_panel;
  tm = (int)(r / width);
  tn = first + (int)(r % width);
}

__global__ void __launch_bounds__(tf::THREADS, 1)
    pairs_tf32_kernel(const __grid_constant__ CUtensorMap map_ahi, const __grid_constant__ CUtensorMap map_alo,
                      const __grid_constant__ CUtensorMap map_bhi, const __grid_constant__ CUtensorMap map_blo,
                      float *__restrict__ out, const double *__restrict__ uu, const double *__restrict__ vv,
                      long long M, long long N, int KB, long long ldo, int kind, int power, int tiles_m,
                      int tiles_n) {
  using namespace tf;
  extern __shared__ unsigned char smem_dyn[];
  // 1024-byte alignment for the 128B-swizzled tiles
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)STAGES * STAGE_BYTES);
  uint64_t *full = bars;                    // [STAGES]  TMA -> MMA
  uint64_t *empty = bars + STAGES;          // [STAGES]  MMA -> TMA
  uint64_t *tmem_full = bars + 2 * STAGES;  // [2]       MMA -> epilogue
  uint64_t *tmem_empty = tmem_full + 2;     // [2]       epilogue -> MMA
  uint32_t *tmem_base_slot = reinterpret_cast<uint32_t *>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long tiles = (long long)tiles_m * tiles_n;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 4);  // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===== TMA producer =======================================================================
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (long long t = blockIdx.x; t < tiles; t += gridDim.x) {
        int tm, tn;
        tile_coords_tf(t, tiles_m, tiles_n, tm, tn);
        const int m0 = tm * BM, n0 = tn * BN;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait_guard(&empty[s], ph ^ 1);
          unsigned char *st = smem + (size_t)s * STAGE_BYTES;
          mbar_expect_tx(&full[s], (uint32_t)STAGE_BYTES);
          tma_load_2d(st, &map_ahi, kb * BK, m0, &full[s]);
          tma_load_2d(st + A_TILE, &map_alo, kb * BK, m0, &full[s]);
          tma_load_2d(st + 2 * A_TILE, &map_bhi, kb * BK, n0, &full[s]);
          tma_load_2d(st + 2 * A_TILE + B_TILE, &map_blo, kb * BK, n0, &full[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer ===========================================================================
    int s = 0;
    uint32_t ph = 0;
    int it = 0;
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
      const int as = it & 1;
      mbar_wait_guard(&tmem_empty[as], (uint32_t)(((it >> 1) & 1) ^ 1));
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait_guard(&full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t st = smem_u32(smem + (size_t)s * STAGE_BYTES);
          const uint64_t ahi = umma_desc_sw128(st), alo = umma_desc_sw128(st + A_TILE);
          const uint64_t bhi = umma_desc_sw128(st + 2 * A_TILE), blo = umma_desc_sw128(st + 2 * A_TILE + B_TILE);
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            const uint64_t adv = (uint64_t)((k * UK * 4) >> 4);  // 32 bytes per K step inside the swizzle atom
            tc_mma_tf32(tmem_d, alo + adv, bhi + adv, IDESC, (kb | k) != 0 ? 1u : 0u);  // small terms first
            tc_mma_tf32(tmem_d, ahi + adv, blo + adv, IDESC, 1u);
            tc_mma_tf32(tmem_d, ahi + adv, bhi + adv, IDESC, 1u);
          }
          tc_commit(&empty[s]);                        // smem stage reusable once these MMAs retire
          if (kb == KB - 1) tc_commit(&tmem_full[as]);  // accumulator complete
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ===== epilogue warps (TMEM lane quarter = warp % 4) ===========================================
    const int q = warp & 3;
    const bool need_norms = (kind != SOAP_KIND_NORMALIZED);
    const bool vec_ok = ((ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    int it = 0;
    for (long long t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
      int tm, tn;
      tile_coords_tf(t, tiles_m, tiles_n, tm, tn);
      const int as = it & 1;
      mbar_wait_guard(&tmem_full[as], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      const long long row = (long long)tm * BM + q * 32 + lane;
      const double ui = (need_norms && row < M) ? uu[row] : 1.0;
      const uint32_t taddr = tmem_base + (uint32_t)(as * BN) + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)(c * 32), v);
        const long long col0 = (long long)tn * BN + c * 32;
        if (row < M && col0 < N) {
          float r[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float g = __uint_as_float(v[j]);
            if (!need_norms) {
              r[j] = sqrtf(fabsf(2.0f - 2.0f * g));  // SOAPdistanceNormalized in float32 (distances.py:83)
            } else {
              const long long col = col0 + j;
              const double vj = col < N ? vv[col] : 1.0;
              r[j] = (float)soap_epilogue((double)g, ui, vj, kind, power);
            }
          }
          float *dst = out + (size_t)row * ldo + col0;
          if (vec_ok && col0 + 32 <= N) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 o = make_float4(r[j], r[j + 1], r[j + 2], r[j + 3]);
              *reinterpret_cast<float4 *>(dst + j) = o;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col0 + j < N) dst[j] = r[j];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ---- host side ----------------------------------------------------------------------------------
static int make_map(CUtensorMap *map, const float *base, int64_t rows, int Dp, int box_rows) {
  return make_tensor_map_2d(map, base, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, rows, Dp, tf::BK, box_rows,
                            CU_TENSOR_MAP_SWIZZLE_128B);
}

size_t pairs_tf32_scratch(int64_t M, int64_t N, int D) {
  const size_t Dp = align_up((size_t)D, tf::BK);
  return 2 * align_up((size_t)M * Dp * 4, 1024) + 2 * align_up((size_t)N * Dp * 4, 1024) + (size_t)(M + N) * 8 + 2048;
}

int pairs_tf32(const float *A, const float *B, float *out, int64_t M, int64_t N, int D, int64_t ldo, int kind,
               int power, void *scratch, cudaStream_t s) {
  SOAP_REQUIRE(out != nullptr, "soap_pairs: the tensor-core path writes the full matrix (out == NULL)");
  const int Dp = (int)align_up((size_t)D, tf::BK);
  unsigned char *p = reinterpret_cast<unsigned char *>(align_up(reinterpret_cast<uintptr_t>(scratch), 1024));
  const bool same = (A == B && M == N);
  float *ahi = reinterpret_cast<float *>(p);
  p += align_up((size_t)M * Dp * 4, 1024);
  float *alo = reinterpret_cast<float *>(p);
  p += align_up((size_t)M * Dp * 4, 1024);
  float *bhi = ahi, *blo = alo;
  if (!same) {
    bhi = reinterpret_cast<float *>(p);
    p += align_up((size_t)N * Dp * 4, 1024);
    blo = reinterpret_cast<float *>(p);
    p += align_up((size_t)N * Dp * 4, 1024);
  }
  double *uu = reinterpret_cast<double *>(p);
  double *vv = same ? uu : uu + M;
  split_tf32_kernel<<<(unsigned)((M + 7) / 8), 256, 0, s>>>(A, ahi, alo, uu, M, D, Dp);
  int rc = check_launch("split_tf32_kernel");
  if (rc) return rc;
  if (!same) {
    split_tf32_kernel<<<(unsigned)((N + 7) / 8), 256, 0, s>>>(B, bhi, blo, vv, N, D, Dp);
    rc = check_launch("split_tf32_kernel");
    if (rc) return rc;
  }
  CUtensorMap m_ahi, m_alo, m_bhi, m_blo;
  if ((rc = make_map(&m_ahi, ahi, M, Dp, tf::BM))) return rc;
  if ((rc = make_map(&m_alo, alo, M, Dp, tf::BM))) return rc;
  if ((rc = make_map(&m_bhi, bhi, N, Dp, tf::BN))) return rc;
  if ((rc = make_map(&m_blo, blo, N, Dp, tf::BN))) return rc;
  const int tiles_m = (int)((M + tf::BM - 1) / tf::BM), tiles_n = (int)((N + tf::BN - 1) / tf::BN);
  const long long tiles = (long long)tiles_m * tiles_n;
  const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  SOAP_CUDA(cudaFuncSetAttribute(pairs_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tf::SMEM_BYTES));
  ProfScope prof(SOAP_PROF_PAIRS, s);
  pairs_tf32_kernel<<<grid, tf::THREADS, tf::SMEM_BYTES, s>>>(m_ahi, m_alo, m_bhi, m_blo, out, uu, vv, M, N, Dp / tf::BK,
                                                             ldo, kind, power, tiles_m, tiles_n);
  return check_launch("pairs_tf32_kernel");
}

}  // namespace soapb
