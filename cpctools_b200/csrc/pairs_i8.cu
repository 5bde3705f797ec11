// pairs_i8.cu -- K5: the all-pairs SOAP distance matrix on the tcgen05 INT8 tensor cores with EXACT
// accumulation (Ozaki-style digit splitting), the distance as the GEMM epilogue.
//
// Reference semantics: getDistanceBetween (src/SOAPify/classify.py:96-117) with one of the distances
// of src/SOAPify/distances.py; output dtype = data.dtype (classify.py:113).
//
// Why not 3xTF32: the TMEM accumulator of tcgen05.mma.kind::tf32 adds every K=8 slice with fp32
// round-toward-zero, a one-sided bias of ~0.5 ulp per slice; over K = 576 x 3 passes it reaches
// 6e-6 on the kernel value (measured, pairs_tf32.cu), above the 1e-6 tolerance of the fp32 path.
// Integer accumulation has no rounding at all:
//   every row is scaled by the power of two sigma_i >= max|x_i| / (126.5 * 2^(8(ND-1))) and rounded to an integer
//   X = sum_a d_a 2^(8(ND-1-a)) with ND balanced base-256 digits d_a in [-128,127] (int8 planes);
//   x.y = sigma_i sigma_j sum_{a,b} 2^(8(2ND-2-a-b)) (d_a . d_b); the digit products d_a.d_b are
//   int8 x int8 -> int32 tensor-core GEMMs, EXACT (|sum| <= 6 x 576 x 2^14 << 2^31); pairs with
//   a+b > ND-1 are dropped (relative weight <= 2^(-8 ND)), pairs of equal a+b share one accumulator.
//   ND = 3 (24-bit digits) for float32 data: 6 passes, error ~2^-24 of max|x|max|y| per product;
//   ND = 6 (48-bit digits) for float64 data: 21 passes, error ~2^-48: fp64 accuracy at >4x the
//   throughput of the FP64 (DMMA) tensor path.
//
// One persistent CTA per SM, warp specialised like pairs_tf32.cu: warp 0 TMA producer (2 tensor maps
// over the stacked digit planes, swizzled K-major tiles), warp 1 issues tcgen05.mma.kind::i8
// (M128 x BN x K32 per instruction, ND(ND+1)/2 passes per K step into ND TMEM accumulators),
// warps 2..9 drain TMEM (tcgen05.ld), recombine the digit sums, apply the distance and store.
#include "tc05.cuh"

#include <stdlib.h>
#include <string.h>

#include <type_traits>
#include <vector>

namespace soapb {

constexpr int kI8BM = 128;
// epilogue warps per CTA (template parameter EW): two per TMEM lane quarter for fp32; three for fp64, whose
// long fp64 dependency chains profit from more warps (A/B: 7.78 -> 7.56 ms; fp32 1.83 -> 1.85)
// bytes per staged row: dense, the 16-byte unit c of row r sits at unit c ^ ((r >> 1) & 3) -- conflict free both for
// the writes (a lane = a row, all lanes the same c) and for the reads (a quarter warp = 2 rows x 4 units)
constexpr int kI8StagePitch = 64;

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
// exact int64 -> double for |x| < 2^51 without the slow I2F.F64.S64: add the integer to the bit pattern of
// 1.5 * 2^52 (its ulp is 1) and subtract that constant again
__device__ __forceinline__ double exact_i2d(long long x) {
  return __longlong_as_double(x + 0x4338000000000000ll) - 6755399441055744.0;
}
// shared memory by 32-bit address (a generic pointer into the dynamic array compiles to LD.E / ST.E, whose data return
// through the long-scoreboard path: measured as the largest stall of the epilogue)
__device__ __forceinline__ void sts_f4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ int4 lds_i4(uint32_t addr) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ float4 lds_f4_sync(uint32_t addr) {  // ordered against the st.shared above ("memory")
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// distance from the exact dot g and the reciprocal norms (rr = 1/|x| * 1/|y|) for the kinds that
// normalise inside: SIMPLE mirrors scipy's clipped cosine (distances.py:18,33), KERNEL distances.py:51,66
__device__ __forceinline__ double fast_cosine_distance(double g, double rr, int kind, int power) {
  const double c = g * rr;
  if (kind == SOAP_KIND_SIMPLE) {
    double cosd = 1.0 - c;
    cosd = (cosd != cosd) ? cosd : fmin(fmax(cosd, 0.0), 2.0);
    return sqrt(2.0 - 2.0 * (1.0 - cosd));
  }
  return sqrt(2.0 - 2.0 * ipow(c, power));
}

// fp32 outputs: the exact dot (one fp32 rounding) is normalised in double and 2 - 2k formed with one DFMA
// (the cancellation for k ~ 1 happens in double); clip and square root run in fp32.  For SIMPLE,
// 2 - 2 (1 - clip(1 - c, 0, 2)) == clip(2 - 2c, 0, 4) up to 1e-16; a negative KERNEL argument gives NaN
// like numpy's sqrt.
__device__ __forceinline__ float cosine_distance_f32(float g, double rr, int kind, int power) {
  const double c = (double)g * rr;
  float d2 = (float)fma(-2.0, kind == SOAP_KIND_SIMPLE ? c : ipow(c, power), 2.0);
  if (kind == SOAP_KIND_SIMPLE) d2 = (d2 != d2) ? d2 : fminf(fmaxf(d2, 0.0f), 4.0f);
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(d2));
  return r;
}

// ---- pre-pass: row scale, digit planes (zero padded to Dp), squared row norms ----------------------
// planes[a][row][k] int8; scale[row] = sigma * 2^(8(ND-1)) so that  x.y = scale_i scale_j sum_g S_g 2^(-8g)
template <typename T, int ND>
__global__ void __launch_bounds__(256)
    ozaki_split_kernel(const T *__restrict__ X, int8_t *__restrict__ planes, double *__restrict__ scale,
                       double *__restrict__ norm2, long long rows, int D, int Dp) {
  const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const T *x = X + (size_t)row * D;
  double mx = 0.0, acc = 0.0;
  for (int k = lane; k < D; k += 32) {
    const double v = (double)x[k];
    mx = fmax(mx, fabs(v));
    acc = fma(v, v, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  acc = warp_sum(acc);
  const double cap = 126.5 * (double)(1ull << (8 * (ND - 1)));
  const bool finite = (acc == acc) && (acc < 1.0e308);
  // power-of-two scale >= max|x| / cap: scaling and un-scaling are then exact in every precision
  int sexp = 0;
  if (mx > 0.0 && finite) {
    frexp(mx / cap, &sexp);  // mx/cap = m * 2^sexp, m in [0.5, 1)
  }
  const double sigma = ldexp(1.0, sexp);
  const double inv = 1.0 / sigma;
  if (lane == 0) {
    scale[row] = finite ? sigma * (double)(1ull << (8 * (ND - 1))) : __longlong_as_double(0x7ff8000000000000ll);
    norm2[row] = acc;
  }
  const size_t plane_stride = (size_t)rows * Dp;
  for (int k0 = 4 * lane; k0 < Dp; k0 += 128) {
    int8_t dig[ND][4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = k0 + e;
      long long v = 0;
      if (k < D && finite) v = __double2ll_rn((double)x[k] * inv);
      // balanced base-256 digits, least significant first with carry: v == sum_a dig[a] 256^(ND-1-a)
#pragma unroll
      for (int a = ND - 1; a >= 0; --a) {
        int d = (int)(v & 0xff);
        if (d > 127) d -= 256;
        dig[a][e] = (int8_t)d;
        v = (v - d) >> 8;
      }
    }
#pragma unroll
    for (int a = 0; a < ND; ++a) {
      char4 c = make_char4(dig[a][0], dig[a][1], dig[a][2], dig[a][3]);
      *reinterpret_cast<char4 *>(planes + a * plane_stride + (size_t)row * Dp + k0) = c;
    }
  }
}

__device__ __forceinline__ void tile_coords_i8(long long t, int tiles_m, int tiles_n, int &tm, int &tn) {
  constexpr int PW = 6;
  const long long per_panel = (long long)PW * tiles_m;
  const int panel = (int)(t / per_panel);
  const int first = panel * PW;
  const int width = min(PW, tiles_n - first);
  const long long r = t - (long long)panel * per_panel;
  tm = (int)(r / width);
  tn = first + (int)(r % width);
}

// numpy.argmin / numpy.amin order for values visited in increasing column order: the first NaN wins, then the
// smaller value, then the earlier column
template <typename V>
__device__ __forceinline__ void am_update(V v, int j, V &bv, int &bj) {
  if (bj < 0 || (bv == bv && (v != v || v < bv))) bv = v, bj = j;
}

// The same rule on an integer key for the float32 distances (all >= +0 or NaN): NaN -> 0, otherwise bits + 1, so that
// one unsigned compare decides (strict <: the earlier column keeps a tie, the first NaN stays).  Start: key ~0u.
// Branch free on purpose: a per-group "can this change the minimum?" test in front of the exact rule measured
// SLOWER (19.1 against 17.3 ms for the bench block) -- the divergent branch costs more than the selects.
__device__ __forceinline__ void am_update_key(float v, int j, uint32_t &bk, int &bj) {
  const uint32_t k = (v != v) ? 0u : __float_as_uint(v) + 1u;
  if (k < bk) bk = k, bj = j;
}

constexpr int kI8PanelW = 6;        // tile columns per panel (L2 reuse of the B tiles)
constexpr int kI8MaxPanels = 2048;  // symmetric calls: panel offsets live in shared memory

// AM: also keep, per row and (tile column, epilogue warp), the minimum over the columns and its index
// (am_val / am_idx [slot][row], slot = tn * (EW / 4) + part; rowmin_reduce_kernel folds the slots); `out` may
// then be NULL.  Column row + excl is skipped (the self distance of a row block).
// CG = 2: CTA pairs.  The two CTAs of a cluster (the two SMs of a TPC) compute a 256 x BN tile with ONE
// tcgen05.mma.cta_group::2 per pass, issued by the even CTA: each CTA stages its own 128 rows of A and HALF of the B
// tile (BN / 2 rows), so an MMA reads A 4 kB + B BN/2 * 32 B per CTA instead of A + the whole B tile -- what made the
// N = 80 passes of the fp64 kernel shared-memory-operand bound -- and a stage is smaller (one more fits).  TMA loads
// of both CTAs are counted on the leader's `full` barrier, tcgen05.commit is multicast to the `empty` / `tmem_full`
// barriers of both, and the epilogue warps of both arrive on the leader's `tmem_empty`.
template <int ND, int BN, int BKB, int ACC, int STAGES, bool FAST, bool SYM, int EW, bool COS, bool AM, typename OutT, int CG = 1>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
    pairs_i8_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                    OutT *__restrict__ out, const double *__restrict__ sa, const double *__restrict__ sb,
                    const double *__restrict__ uu, const double *__restrict__ vv, long long M, long long N, int KB,
                    long long ldo, int kind, int power, int tiles_m, int tiles_n, int debug_flags,
                    double *__restrict__ am_val, int *__restrict__ am_idx, long long excl) {
  constexpr int BM = kI8BM;
  constexpr int kI8EpiWarps = EW;
  static_assert(CG == 1 || CG == 2, "CTAs per MMA");
  static_assert(!(SYM && CG == 2), "the symmetric tile walk is built for single CTAs");
  constexpr int B_ROWS = BN / CG;  // rows of the B tile staged by this CTA
  constexpr int A_PLANE = BM * BKB, B_PLANE = B_ROWS * BKB;
  constexpr int STAGE_BYTES = ND * (A_PLANE + B_PLANE);
  constexpr int KS = BKB / 32;  // UMMA K = 32 for 8-bit operands
  // instruction descriptor: D = s32, A = B = signed int8, K-major, N = BN, M = 128 (256 for a CTA pair)
  constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * CG) >> 4) << 24);
  static_assert(ACC * ND * BN <= 512, "accumulators exceed TMEM");  // ACC = accumulator sets (2: epilogue overlaps the next tile)
  static_assert(BN % 16 == 0 && BN <= 256 && (B_ROWS % 8) == 0, "invalid UMMA N");
  const int crank = CG == 2 ? (int)cluster_ctarank() : 0;            // 0 = leader of the pair
  const long long work0 = CG == 2 ? (long long)(blockIdx.x >> 1) : (long long)blockIdx.x;  // first tile of this CTA (pair)
  const long long work_step = CG == 2 ? (long long)(gridDim.x >> 1) : (long long)gridDim.x;

  extern __shared__ unsigned char smem_dyn[];
  unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)STAGES * STAGE_BYTES);
  uint64_t *full = bars, *empty = bars + STAGES;
  uint64_t *tmem_full = bars + 2 * STAGES, *tmem_empty = tmem_full + ACC;
  uint32_t *tmem_base_slot = reinterpret_cast<uint32_t *>(tmem_empty + ACC);
  int *col_exp = reinterpret_cast<int *>(smem + (size_t)STAGES * STAGE_BYTES + 256);  // [kI8EpiWarps][BN] (FAST path)
  double *col_rn = reinterpret_cast<double *>(col_exp + kI8EpiWarps * BN);            // [kI8EpiWarps][BN], cosine kinds only
  // FAST epilogues: per-warp [32 rows][64 + 16 pad bytes] tile that turns a chunk of the accumulator
  // (lane = row) into row-contiguous 64-byte stores
  unsigned char *epi_stage = reinterpret_cast<unsigned char *>(col_rn + kI8EpiWarps * BN);
  // symmetric calls: first tile of every panel in the walk below (after the staging tiles)
  long long *panel_off = reinterpret_cast<long long *>(epi_stage + (FAST ? kI8EpiWarps * 32 * kI8StagePitch : 0));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // symmetric mode (data is spectra): only tiles that reach the upper triangle are computed, and every value is
  // also written to its mirrored position.  Panels of 6 tile columns are walked row by row (L2 reuse) down to
  // the last tile row that still touches the upper triangle; (tm, tn) follows from the tile number through the
  // per-panel offsets -- no tile list, no host work per call.  The few tiles of a panel's last rows that lie
  // wholly below the diagonal are skipped by every role.
  constexpr bool symmetric = SYM;
  const int n_panels = (tiles_n + kI8PanelW - 1) / kI8PanelW;
  auto panel_rows = [&](int p) {  // tile rows tm with tm * BM < (last column of the panel + 1) * BN
    const long long last_col = min((long long)(p + 1) * kI8PanelW, (long long)tiles_n) * BN;
    return (int)min((long long)tiles_m, (last_col + BM - 1) / BM);
  };
  if (symmetric) {
    if (threadIdx.x == 0) {
      long long acc = 0;
      for (int p = 0; p < n_panels; ++p) {
        panel_off[p] = acc;
        acc += (long long)panel_rows(p) * min(kI8PanelW, tiles_n - p * kI8PanelW);
      }
      panel_off[n_panels] = acc;
    }
    __syncthreads();
  }
  const long long tiles = symmetric ? panel_off[n_panels] : (long long)tiles_m * tiles_n;
  // returns false for a tile that needs no work (symmetric mode: wholly below the diagonal)
  auto coords = [&](long long t, int &tm, int &tn) -> bool {
    if (symmetric) {
      int lo = 0, hi = n_panels - 1;
      while (lo < hi) {  // last panel whose first tile is <= t
        const int mid = (lo + hi + 1) >> 1;
        if (panel_off[mid] <= t) lo = mid; else hi = mid - 1;
      }
      const int first = lo * kI8PanelW, width = min(kI8PanelW, tiles_n - first);
      const long long r = t - panel_off[lo];
      tm = (int)(r / width);
      tn = first + (int)(r % width);
      return (long long)(tn + 1) * BN > (long long)tm * BM;
    }
    tile_coords_i8(t, tiles_m, tiles_n, tm, tn);  // CG == 2: tiles_m counts 256-row tiles
    tm = tm * CG + crank;
    return true;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < ACC; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], kI8EpiWarps * CG);
    }
    fence_mbar_init();
  }
  if constexpr (CG == 2) cluster_sync_all();  // the peer's barriers exist before anything is sent to them
  if (warp == 1) {
    if constexpr (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "n"(512)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)), "n"(512)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp == 0) {
    // ===== TMA producer: digit plane a of the operand lives at rows [a*rows, (a+1)*rows) ==========
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      for (long long t = work0; t < tiles; t += work_step) {
        int tm, tn;
        if (!coords(t, tm, tn)) continue;
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait_guard(&empty[s], ph ^ 1);
          unsigned char *st = smem + (size_t)s * STAGE_BYTES;
          if constexpr (CG == 2) {
            // the leader's barrier counts the bytes of both CTAs (a stage is released to both by one multicast
            // commit, so the peer's loads of this round cannot land before the leader's barrier is in this phase)
            if (crank == 0) mbar_expect_tx(&full[s], 2u * (uint32_t)STAGE_BYTES);
#pragma unroll
            for (int a = 0; a < ND; ++a) {
              tma_load_2d_pair(st + a * A_PLANE, &map_a, kb * BKB, (int)(a * M + (long long)tm * BM), &full[s]);
              tma_load_2d_pair(st + ND * A_PLANE + a * B_PLANE, &map_b, kb * BKB,
                               (int)(a * N + (long long)tn * BN + crank * B_ROWS), &full[s]);
            }
          } else {
            mbar_expect_tx(&full[s], (uint32_t)STAGE_BYTES);
#pragma unroll
            for (int a = 0; a < ND; ++a) {
              tma_load_2d(st + a * A_PLANE, &map_a, kb * BKB, (int)(a * M + (long long)tm * BM), &full[s]);
              tma_load_2d(st + ND * A_PLANE + a * B_PLANE, &map_b, kb * BKB, (int)(a * N + (long long)tn * BN), &full[s]);
            }
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer ===========================================================================
    int s = 0;
    uint32_t ph = 0;
    int it = 0;
    for (long long t = work0; t < tiles && crank == 0; t += work_step) {  // (a pair's MMAs are issued by its leader)
      {
        int tm, tn;
        if (!coords(t, tm, tn)) continue;
      }
      const int as = it % ACC;
      const int it_now = it++;
      mbar_wait_guard(&tmem_empty[as], (uint32_t)(((it_now / ACC) & 1) ^ 1));  // epilogue has drained this accumulator set
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + (uint32_t)(as * ND * BN);
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait_guard(&full[s], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t st = smem_u32(smem + (size_t)s * STAGE_BYTES);
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            const uint64_t adv = (uint64_t)((ks * 32) >> 4);
#pragma unroll
            for (int a = 0; a < ND; ++a) {
              const uint32_t pa = st + a * A_PLANE;
              const uint64_t da = (BKB == 128 ? umma_desc_sw128(pa) : BKB == 64 ? umma_desc_sw64(pa) : umma_desc_sw32(pa)) + adv;
#pragma unroll
              for (int b = 0; b < ND - a; ++b) {
                const uint32_t pb = st + ND * A_PLANE + b * B_PLANE;
                const uint64_t db = (BKB == 128 ? umma_desc_sw128(pb) : BKB == 64 ? umma_desc_sw64(pb) : umma_desc_sw32(pb)) + adv;
                // digit pairs of equal weight a+b share accumulator g = a+b; (0, g) is its first pass
                if (!(debug_flags & 2) || (a | b | ks) == 0) {
                  if constexpr (CG == 2)
                    tc_mma_i8_pair(tmem_acc + (uint32_t)((a + b) * BN), da, db, IDESC, (kb | ks | a) != 0 ? 1u : 0u);
                  else
                    tc_mma_i8(tmem_acc + (uint32_t)((a + b) * BN), da, db, IDESC, (kb | ks | a) != 0 ? 1u : 0u);
                }
              }
            }
          }
          if constexpr (CG == 2) {
            tc_commit_pair(&empty[s]);
            if (kb == KB - 1) tc_commit_pair(&tmem_full[as]);
          } else {
            tc_commit(&empty[s]);
            if (kb == KB - 1) tc_commit(&tmem_full[as]);
          }
        }
        __syncwarp();
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ===== epilogue warps (TMEM lane quarter = warp % 4) ===========================================
    const int q = warp & 3;               // TMEM lane quarter this warp may read
    const int part = (warp - 2) >> 2;      // column chunks are interleaved over the warps of a quarter
    constexpr int PARTS = kI8EpiWarps / 4;
    const bool need_norms = (kind != SOAP_KIND_NORMALIZED);
    constexpr int VEC = 16 / (int)sizeof(OutT);
    const bool vec_ok = ((ldo % VEC) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    int it = -1;
    for (long long t = work0; t < tiles; t += work_step) {
      int tm, tn;
      if (!coords(t, tm, tn)) continue;
      ++it;
      const int as = it % ACC;
      if constexpr (FAST) {
        // SOAPdistanceNormalized (distances.py:83) with a lean, kind-specific epilogue.  fp32: entirely in
        // fp32 -- the scales are powers of two (exact), tf carries ONE rounding, 2-2g is exact for g in
        // [1/2, 2] (Sterbenz), the square root is the 1-ulp hardware approximation.  Global reads (row /
        // column scale exponents) are issued before the accumulator wait; column exponents sit in a
        // per-warp shared array.
        const long long row = (long long)tm * BM + q * 32 + lane;
        const bool row_ok = row < M;
        const int ei = row_ok ? ((__double2hiint(sa[row]) >> 20) & 0x7ff) : 0x7ff;
        int *my_exp = col_exp + (warp - 2) * BN;
        double *my_rn = col_rn + (warp - 2) * BN;
        constexpr bool cosine = COS;  // SIMPLE / KERNEL (normalise inside) vs SOAPdistanceNormalized
        const double ri = (cosine && row_ok) ? rsqrt(uu[row]) : 1.0;  // 1/|x_i|; a zero vector gives inf -> NaN like 0/0
        // When every scale exponent of this warp's rows and of the tile's columns is moderate (always, for
        // data of ordinary magnitude) the power-of-two factors are staged as floats: c_j = -2 * 2^(e_j - 1023),
        // and an output costs  3 I2F + 3 FFMA + 1 FMUL + 1 MUFU.  Otherwise the exponents are staged as
        // integers and combined with clamping / NaN propagation per output (general path).
        bool lean = false;
        if constexpr (!COS) {
          constexpr int kSpan = ND == 3 ? 60 : 480;  // the product of two factors stays a normal number
          bool ok = !row_ok || (ei >= 1023 - kSpan && ei <= 1023 + kSpan);
          int ejs[(BN + 31) / 32];
#pragma unroll
          for (int r = 0; r < (BN + 31) / 32; ++r) {
            const int cidx = lane + 32 * r;
            const long long col = (long long)tn * BN + cidx;
            ejs[r] = (cidx < BN && col < N) ? ((__double2hiint(sb[col]) >> 20) & 0x7ff) : 1023;
            ok = ok && (ejs[r] >= 1023 - kSpan && ejs[r] <= 1023 + kSpan);
          }
          lean = __all_sync(0xffffffffu, ok);
          if (lean) {
#pragma unroll
            for (int r = 0; r < (BN + 31) / 32; ++r) {
              const int cidx = lane + 32 * r;
              if (cidx < BN) {  // c_j = -2 * 2^(e_j - 1023) = -2^(e_j - 1022)
                if constexpr (ND == 3)
                  my_exp[cidx] = (int)(0x80000000u | (uint32_t)((ejs[r] - 1023 + 128) << 23));
                else
                  my_rn[cidx] = __hiloint2double((int)(0x80000000u | (uint32_t)((ejs[r] + 1) << 20)), 0);
              }
            }
          }
        }
        if (!lean) {
          for (int cidx = lane; cidx < BN; cidx += 32) {
            const long long col = (long long)tn * BN + cidx;
            my_exp[cidx] = col < N ? ((__double2hiint(sb[col]) >> 20) & 0x7ff) : 0x7ff;
            if (cosine) my_rn[cidx] = col < N ? rsqrt(vv[col]) : 1.0;
          }
        }
        __syncwarp();
        mbar_wait_guard(&tmem_full[as], (uint32_t)((it / ACC) & 1));
        tc_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t)(as * ND * BN) + ((uint32_t)(q * 32) << 16);
        if constexpr (ND == 3) {
          const bool vec4 = ((ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
          const float prow = __int_as_float((min(max(ei - 1023, -126), 127) + 127) << 23);  // used when `lean`
          uint32_t best_k = ~0u;  // AM: running minimum (as a key, am_update_key) of this lane's row over this warp's chunks
          int best_j = -1;
          // AM: no per-output column tests in a tile that has all its columns and does not meet the excluded diagonal
          const long long ex_lo = (long long)tm * BM + excl;
          const bool am_plain = (long long)(tn + 1) * BN <= N && (ex_lo + BM <= (long long)tn * BN || ex_lo >= (long long)(tn + 1) * BN);
          // one chunk loop, three compile-time flavours of the per-output arithmetic:
          // 0 = lean SOAPdistanceNormalized, 1 = general SOAPdistanceNormalized, 2 = cosine kinds
          // interior tile: no row / column bound checks on the stores (all but the last tile row / column)
          const bool interior = vec4 && (long long)(tm + 1) * BM <= M && (long long)(tn + 1) * BN <= N;
          const uint32_t exp_a = smem_u32(my_exp);
          const uint32_t stg_a = smem_u32(epi_stage) + (uint32_t)((warp - 2) * (32 * kI8StagePitch));
          // staged chunk -> global: a lane takes 16 bytes of row 8 m + (lane >> 2); 4 lanes write the 64 contiguous
          // bytes of a row (a quarter warp touches two rows: fewest cache lines per store pass)
          const int srow = lane >> 2, cq = (lane & 3) * 4;
          const uint32_t rd_a = stg_a + (uint32_t)(srow * kI8StagePitch + (((lane & 3) ^ ((srow >> 1) & 3)) << 4));
          const uint32_t wr_sw = (uint32_t)((lane >> 1) & 3);
          float *orow[4];
#pragma unroll
          for (int m = 0; m < 4; ++m)
            orow[m] = reinterpret_cast<float *>(out) + (size_t)((long long)tm * BM + q * 32 + 8 * m + srow) * ldo + (long long)tn * BN + cq;
          auto chunk_loop = [&](auto mode_tag) {
          constexpr int MODE = decltype(mode_tag)::value;
#pragma unroll 1
          for (int c = part; c < ((debug_flags & 1) ? 0 : BN / 16); c += PARTS) {
            uint32_t v[ND][16];
#pragma unroll
            for (int g = 0; g < ND; ++g) tmem_ld16(taddr + (uint32_t)(g * BN + c * 16), v[g]);
            int ej[16];  // lean: the bits of c_j = -2 * 2^(e_j - 1023) as floats; otherwise biased exponents
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              const int4 e4 = lds_i4(exp_a + (uint32_t)(c * 16 + j) * 4u);
              ej[j] = e4.x; ej[j + 1] = e4.y; ej[j + 2] = e4.z; ej[j + 3] = e4.w;
            }
            tmem_ld_wait();
            const long long col0 = (long long)tn * BN + c * 16;
            // a lane owns a ROW of the accumulator: storing from registers sends 32 requests of 16 bytes per
            // instruction to L2, one per clock (measured: 0.55 of 2.1 ms exposed).  The chunk goes through a
            // padded shared-memory tile instead and leaves as 64 contiguous bytes per row (4 lanes each).
            // (A TMA tile store of the same chunk measured slower: the proxy fence sits on the critical path.)
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float o[4];
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                const int jj = j + k4;
                const float tf = fmaf(fmaf((float)(int)v[2][jj], 1.0f / 256.0f, (float)(int)v[1][jj]), 1.0f / 256.0f,
                                      (float)(int)v[0][jj]);
                // (.ftz: the argument is 0 or at least ~2^-46, never subnormal; plain sqrt.approx wraps the MUFU in a
                // scale / compare / select sequence for subnormals, 4 of the ~30 instructions per output)
                if constexpr (MODE == 0) {
                  // 2 - 2 g with g = tf * 2^(e_i + e_j): the scaling is exact, so this is the general path's value
                  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(o[k4]) : "f"(fabsf(fmaf(prow * __int_as_float(ej[jj]), tf, 2.0f))));
                } else {
                  int k = ei + ej[jj] - 2 * 1023;
                  k = k < -126 ? -126 : (k > 127 ? 127 : k);
                  const float pw = (ei == 0x7ff || ej[jj] == 0x7ff) ? __int_as_float(0x7fc00000) : __int_as_float((k + 127) << 23);
                  if constexpr (MODE == 2) {
                    o[k4] = cosine_distance_f32(tf * pw, ri * my_rn[c * 16 + jj], kind, power);
                  } else {
                    asm("sqrt.approx.f32 %0, %1;" : "=f"(o[k4]) : "f"(fabsf(fmaf(-2.0f, tf * pw, 2.0f))));
                  }
                }
              }
              if constexpr (AM) {
                if (am_plain) {
#pragma unroll
                  for (int k4 = 0; k4 < 4; ++k4) am_update_key(o[k4], (int)col0 + j + k4, best_k, best_j);
                } else {
#pragma unroll
                  for (int k4 = 0; k4 < 4; ++k4) {
                    const long long col = col0 + j + k4;
                    if (col < N && col != row + excl) am_update_key(o[k4], (int)col, best_k, best_j);
                  }
                }
              }
              if (!AM || out != nullptr)
                sts_f4(stg_a + (uint32_t)(lane * kI8StagePitch) + ((((uint32_t)j >> 2) ^ wr_sw) << 4), o[0], o[1], o[2], o[3]);
              if (symmetric && row_ok) {  // out[col][row]: the 32 lanes of a warp write 32 consecutive rows
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4)
                  if (col0 + j + k4 < N) reinterpret_cast<float *>(out)[(size_t)(col0 + j + k4) * ldo + row] = o[k4];
              }
            }
            __syncwarp();
            if (!(debug_flags & 4) && (!AM || out != nullptr)) {
              if (interior) {
                float4 val[4];
#pragma unroll
                for (int m = 0; m < 4; ++m) val[m] = lds_f4_sync(rd_a + (uint32_t)(8 * m * kI8StagePitch));
#pragma unroll
                for (int m = 0; m < 4; ++m) *reinterpret_cast<float4 *>(orow[m] + c * 16) = val[m];
              } else {
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                  const long long grow = (long long)tm * BM + q * 32 + 8 * m + srow;
                  const float4 val = lds_f4_sync(rd_a + (uint32_t)(8 * m * kI8StagePitch));
                  if (grow < M) {
                    float *dst = orow[m] + c * 16;
                    if (vec4 && col0 + cq + 4 <= N) {
                      *reinterpret_cast<float4 *>(dst) = val;
                    } else {
                      if (col0 + cq + 0 < N) dst[0] = val.x;
                      if (col0 + cq + 1 < N) dst[1] = val.y;
                      if (col0 + cq + 2 < N) dst[2] = val.z;
                      if (col0 + cq + 3 < N) dst[3] = val.w;
                    }
                  }
                }
              }
            }
            __syncwarp();
          }
          };
          if constexpr (COS) {
            chunk_loop(std::integral_constant<int, 2>{});
          } else {
            if (lean)
              chunk_loop(std::integral_constant<int, 0>{});
            else
              chunk_loop(std::integral_constant<int, 1>{});
          }
          if constexpr (AM) {
            if (row_ok) {
              const size_t slot = (size_t)tn * PARTS + part;
              am_val[slot * M + row] = best_k == 0u ? (double)__int_as_float(0x7fc00000) : (double)__uint_as_float(best_k - 1u);
              am_idx[slot * M + row] = best_j;
            }
          }
        } else {
          // fp64 SOAPdistanceNormalized: the six digit sums are folded into two exact 64-bit integers,
          // converted once each; the power-of-two scaling is one exact multiply; IEEE sqrt
          const bool vec2 = ((ldo & 1) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
          constexpr double kHi = 1.0 / (double)(1ull << (8 * (ND / 2 - 1)));
          constexpr double kLo = 1.0 / (double)(1ull << (8 * (ND - 1)));
          const double prow = __hiloint2double(ei << 20, 0);  // 2^(e_i - 1023), only used when `lean`
          double best_v = 0.0;
          int best_j = -1;
          // interior tile: no row / column bound checks on the stores; the per-warp column tables are read by 32-bit
          // shared addresses (generic pointers compile to LD.E: long-scoreboard latency in front of every DMUL)
          const bool interior = vec2 && (long long)(tm + 1) * BM <= M && (long long)(tn + 1) * BN <= N;
          const uint32_t exp_a = smem_u32(my_exp), rn_a = smem_u32(my_rn);
#pragma unroll 1
          for (int c = part; c < ((debug_flags & 1) ? 0 : BN / 8); c += PARTS) {
            uint32_t v[ND][8];
#pragma unroll
            for (int g = 0; g < ND; ++g) tmem_ld8(taddr + (uint32_t)(g * BN + c * 8), v[g]);
            int ej[8];
            double rn[8];
#pragma unroll
            for (int j = 0; j < 8; j += 4) {
              const int4 e4 = lds_i4(exp_a + (uint32_t)(c * 8 + j) * 4u);
              ej[j] = e4.x; ej[j + 1] = e4.y; ej[j + 2] = e4.z; ej[j + 3] = e4.w;
            }
            if (lean || cosine) {
#pragma unroll
              for (int j = 0; j < 8; j += 2) {
                const double2 r2 = lds_d2(rn_a + (uint32_t)(c * 8 + j) * 8u);
                rn[j] = r2.x; rn[j + 1] = r2.y;
              }
            }
            tmem_ld_wait();
            const long long col0 = (long long)tn * BN + c * 8;
            // (staging this chunk through shared memory like the fp32 one measured slower: the fp64 epilogue is
            // bound by its arithmetic, the stores cost 0.3 of 7.8 ms.  A branch-free square root -- hardware
            // reciprocal square root + two coupled Newton steps instead of the IEEE routine, whose special-case
            // branch keeps the eight outputs of a chunk from being interleaved -- measured 63.5 against 63.7 ms for
            // the bench block: not worth giving up the correctly rounded result)
            double *dst = reinterpret_cast<double *>(out) + (size_t)row * ldo + col0;
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
              double o[2];
#pragma unroll
              for (int k2 = 0; k2 < 2; ++k2) {
                const int jj = j + k2;
                long long hi = 0, lo = 0;
#pragma unroll
                for (int g = 0; g < ND / 2; ++g) hi = hi * 256 + (long long)(int)v[g][jj];
#pragma unroll
                for (int g = ND / 2; g < ND; ++g) lo = lo * 256 + (long long)(int)v[g][jj];
                const double tsum = fma(exact_i2d(lo), kLo, exact_i2d(hi) * kHi);
                if (lean) {  // 2 - 2 g with g = tsum * 2^(e_i + e_j): exact scaling, same value as below
                  o[k2] = sqrt(fabs(fma(prow * rn[jj], tsum, 2.0)));
                  continue;
                }
                int k = ei + ej[jj] - 2 * 1023;
                k = k < -1022 ? -1022 : (k > 1023 ? 1023 : k);
                const double pw = (ei == 0x7ff || ej[jj] == 0x7ff) ? __longlong_as_double(0x7ff8000000000000ll)
                                                                   : __hiloint2double((k + 1023) << 20, 0);
                o[k2] = cosine ? fast_cosine_distance(tsum * pw, ri * rn[jj], kind, power)
                               : sqrt(fabs(fma(-2.0, tsum * pw, 2.0)));
              }
              if constexpr (AM) {
                // (a per-tile path without the column tests, as in the fp32 loop, measured 1.5 % slower here, three runs)
#pragma unroll
                for (int k2 = 0; k2 < 2; ++k2) {
                  const long long col = col0 + j + k2;
                  if (col < N && col != row + excl) am_update(o[k2], (int)col, best_v, best_j);
                }
              }
              if (interior && !symmetric && (!AM || out != nullptr)) {
                *reinterpret_cast<double2 *>(dst + j) = make_double2(o[0], o[1]);
              } else if (row_ok && (!AM || out != nullptr)) {
                if (vec2 && col0 + j + 2 <= N) {
                  *reinterpret_cast<double2 *>(dst + j) = make_double2(o[0], o[1]);
                } else {
#pragma unroll
                  for (int k2 = 0; k2 < 2; ++k2)
                    if (col0 + j + k2 < N) dst[j + k2] = o[k2];
                }
                if (symmetric) {
#pragma unroll
                  for (int k2 = 0; k2 < 2; ++k2)
                    if (col0 + j + k2 < N) reinterpret_cast<double *>(out)[(size_t)(col0 + j + k2) * ldo + row] = o[k2];
                }
              }
            }
          }
          if constexpr (AM) {
            if (row_ok) {
              const size_t slot = (size_t)tn * PARTS + part;
              am_val[slot * M + row] = best_v;
              am_idx[slot * M + row] = best_j;
            }
          }
        }
        // (signalling tmem_empty right after this warp's LAST TMEM read, so that the next tile's MMAs overlap the
        // arithmetic and stores of the last chunk, measured SLOWER: fp32 block 15.4 against 15.1 ms, fp64 68.2 against
        // 66.1 -- the kernel runs under the power cap, overlap buys no time there)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 2) mbar_arrive_leader(&tmem_empty[as]); else mbar_arrive(&tmem_empty[as]);
        }
        continue;
      }
      mbar_wait_guard(&tmem_full[as], (uint32_t)((it / ACC) & 1));
      tc_fence_after();
      const long long row = (long long)tm * BM + q * 32 + lane;
      const bool row_ok = row < M;
      const double si = row_ok ? sa[row] : 0.0;
      const int ei = (__double2hiint(si) >> 20) & 0x7ff;  // biased exponent of the (power-of-two) row scale
      const double ui = (need_norms && row_ok) ? uu[row] : 1.0;
      const uint32_t taddr = tmem_base + (uint32_t)(as * ND * BN) + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c = part; c < ((debug_flags & 1) ? 0 : BN / 16); c += PARTS) {
        uint32_t v[ND][16];
#pragma unroll
        for (int g = 0; g < ND; ++g) tmem_ld16(taddr + (uint32_t)(g * BN + c * 16), v[g]);
        const long long col0 = (long long)tn * BN + c * 16;
        const bool full_chunk = col0 + 16 <= N;
        // column scales of this chunk (same 128 bytes for every lane: broadcast loads, L1 resident)
        double sbj[16];
        if (full_chunk) {
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            const double2 two = *reinterpret_cast<const double2 *>(sb + col0 + j);  // col0 is a multiple of 16
            sbj[j] = two.x;
            sbj[j + 1] = two.y;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) sbj[j] = (col0 + j < N) ? sb[col0 + j] : 0.0;
        }
        tmem_ld_wait();
        if (row_ok && col0 < N) {
          OutT r[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const long long col = col0 + j;
            const bool col_ok = col < N;
            // x.y = scale_i scale_j sum_g S_g 2^(-8g): the digit sums are exact integers, the scales
            // exact powers of two
            double tsum;
            if constexpr (ND == 3) {
              // |S0| < 2^24 is exact in fp32; the two lower digit sums only need ~16 and ~8 bits
              const float tf = fmaf(fmaf((float)(int)v[2][j], 1.0f / 256.0f, (float)(int)v[1][j]), 1.0f / 256.0f,
                                    (float)(int)v[0][j]);
              tsum = (double)tf;
            } else {
              // ND == 6: two exact 64-bit integers (three digit sums each), two conversions
              long long hi = 0, lo = 0;
#pragma unroll
              for (int g = 0; g < ND / 2; ++g) hi = hi * 256 + (long long)(int)v[g][j];
#pragma unroll
              for (int g = ND / 2; g < ND; ++g) lo = lo * 256 + (long long)(int)v[g][j];
              constexpr double kHi = 1.0 / (double)(1ull << (8 * (ND / 2 - 1)));
              constexpr double kLo = 1.0 / (double)(1ull << (8 * (ND - 1)));
              tsum = fma((double)lo, kLo, (double)hi * kHi);
            }
            const double gdot = (si * sbj[j]) * tsum;
            r[j] = (OutT)soap_epilogue(gdot, ui, (need_norms && col_ok) ? vv[col] : 1.0, kind, power);
          }
          OutT *dst = out + (size_t)row * ldo + col0;
          if (vec_ok && full_chunk) {
#pragma unroll
            for (int j = 0; j < 16; j += VEC) *reinterpret_cast<int4 *>(dst + j) = *reinterpret_cast<const int4 *>(&r[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (col0 + j < N) dst[j] = r[j];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
          if constexpr (CG == 2) mbar_arrive_leader(&tmem_empty[as]); else mbar_arrive(&tmem_empty[as]);
        }
    }
  }

  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();  // (pair: the leader's MMAs also write the peer's TMEM)
  if (warp == 1) {
    tc_fence_after();
    if constexpr (CG == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

// folds the per-(tile column, epilogue warp) minima of a row: numpy's order (first NaN, smaller value, smaller index)
__global__ void __launch_bounds__(256)
    rowmin_reduce_kernel(const double *__restrict__ val, const int *__restrict__ idx, long long slots, long long M,
                         int32_t *__restrict__ argmin_out, double *__restrict__ min_out) {
  const long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= M) return;
  double bv = 0.0;
  int bj = -1;
  for (long long sl = 0; sl < slots; ++sl) {
    const int j = idx[sl * M + row];
    if (j < 0) continue;
    const double v = val[sl * M + row];
    const bool vn = v != v, bn = bv != bv;
    bool take;
    if (bj < 0) take = true;
    else if (vn || bn) take = vn && (!bn || j < bj);
    else take = v < bv || (v == bv && j < bj);
    if (take) bv = v, bj = j;
  }
  if (argmin_out) argmin_out[row] = bj;
  if (min_out) min_out[row] = bj < 0 ? __longlong_as_double(0x7ff8000000000000ll) : bv;
}

// ---- host side ----------------------------------------------------------------------------------
template <int ND, int BKB>
static size_t i8_scratch_bytes(int64_t M, int64_t N, int D) {
  const size_t Dp = align_up((size_t)D, BKB);
  return align_up((size_t)ND * M * Dp, 1024) + align_up((size_t)ND * N * Dp, 1024) + (size_t)(M + N) * 16 + 4096;
}
// per-(tile column, epilogue warp) row minima of the AM variants: BN >= 80, at most 3 warps per TMEM quarter
static size_t i8_rowmin_bytes(int64_t M, int64_t N) { return (size_t)((N + 79) / 80) * 3 * (size_t)M * 12 + 256; }

struct I8RowMin {
  int32_t *argmin_out;
  double *min_out;
  long long excl;
};

template <typename InT, int ND, int BN, int BKB, int ACC, int STAGES, bool FAST, int EW, bool COS, typename OutT, int CG = 1>
static int launch_pairs_i8(const InT *A, const InT *B, OutT *out, int64_t M, int64_t N, int D, int64_t ldo, int kind,
                           int power, void *scratch, cudaStream_t s, const I8RowMin *am = nullptr) {
  SOAP_REQUIRE(out != nullptr || am != nullptr, "soap_pairs: the tensor-core path writes the full matrix (out == NULL)");
  SOAP_REQUIRE(am == nullptr || FAST, "soap_pairs: row minima on the tensor-core path exist for the three SOAP distances only");
  SOAP_REQUIRE((int64_t)ND * M < 2147483647LL && (int64_t)ND * N < 2147483647LL, "soap_pairs: too many rows for one call");
  const int Dp = (int)align_up((size_t)D, BKB);
  unsigned char *p = reinterpret_cast<unsigned char *>(align_up(reinterpret_cast<uintptr_t>(scratch), 1024));
  const bool same = (static_cast<const void *>(A) == static_cast<const void *>(B) && M == N);
  int8_t *pa = reinterpret_cast<int8_t *>(p);
  p += align_up((size_t)ND * M * Dp, 1024);
  int8_t *pb = pa;
  if (!same) {
    pb = reinterpret_cast<int8_t *>(p);
    p += align_up((size_t)ND * N * Dp, 1024);
  }
  double *sa = reinterpret_cast<double *>(p);
  double *ua = sa + M;
  double *sb = same ? sa : ua + M;
  double *ub = same ? ua : sb + N;
  unsigned char *p_after = reinterpret_cast<unsigned char *>(align_up(reinterpret_cast<uintptr_t>((same ? ua + M : ub + N)), 16));
  ozaki_split_kernel<InT, ND><<<(unsigned)((M + 7) / 8), 256, 0, s>>>(A, pa, sa, ua, M, D, Dp);
  int rc = check_launch("ozaki_split_kernel");
  if (rc) return rc;
  if (!same) {
    ozaki_split_kernel<InT, ND><<<(unsigned)((N + 7) / 8), 256, 0, s>>>(B, pb, sb, ub, N, D, Dp);
    if ((rc = check_launch("ozaki_split_kernel"))) return rc;
  }
  const CUtensorMapSwizzle swz = BKB == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : BKB == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
  CUtensorMap map_a, map_b;
  if ((rc = make_tensor_map_2d(&map_a, pa, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, (int64_t)ND * M, Dp, BKB, kI8BM, swz))) return rc;
  if ((rc = make_tensor_map_2d(&map_b, pb, CU_TENSOR_MAP_DATA_TYPE_UINT8, 1, (int64_t)ND * N, Dp, BKB, BN / CG, swz))) return rc;
  // (CTA pairs: a tile is 256 rows, each CTA of the pair stages half of the B tile)
  const int tiles_m = (int)((M + kI8BM * CG - 1) / (kI8BM * CG)), tiles_n = (int)((N + BN - 1) / BN);
  long long tiles = (long long)tiles_m * tiles_n;
  // data is spectra: d(i,j) == d(j,i) bit for bit (exact integer sums, commutative scaling), so only
  // the tiles that contain some j >= i are computed and every value is stored twice
  // (the upper-triangular tile walk is computed on the device from the tile number: no tile list, no host work
  // and no stream synchronisation per call)
  const bool sym = CG == 1 && FAST && same && am == nullptr && (tiles_n + kI8PanelW - 1) / kI8PanelW <= kI8MaxPanels &&
                   getenv("SOAP_I8_NO_SYMMETRY") == nullptr;
  if (sym) tiles = (tiles + 1) / 2 + tiles_m;  // grid sizing only
  int grid = (int)(tiles < num_sms() ? tiles : num_sms());
  // the reciprocal-norm staging only exists for the cosine kinds (it would not fit beside 4 stages)
  const size_t SMEM = 1024 + (size_t)STAGES * ND * (kI8BM + BN / CG) * BKB + 256 + (size_t)EW * BN * 4 +
                      (FAST ? (size_t)EW * (BN * 8 + 32 * kI8StagePitch) : 0) + (sym ? (size_t)(kI8MaxPanels + 1) * 8 : 0);
  SOAP_REQUIRE(SMEM <= (size_t)max_smem_optin(), "soap_pairs: internal: tile configuration exceeds shared memory");
  constexpr int PARTS = EW / 4;
  double *am_val = nullptr;
  int *am_idx = nullptr;
  if (am) {
    am_val = reinterpret_cast<double *>(p_after);
    am_idx = reinterpret_cast<int *>(am_val + (size_t)tiles_n * PARTS * M);
  }
  void (*kern)(CUtensorMap, CUtensorMap, OutT *, const double *, const double *, const double *, const double *, long long,
               long long, int, long long, int, int, int, int, int, double *, int *, long long);
  if constexpr (CG == 2) {
    static_assert(FAST, "CTA pairs are built for the lean epilogues");
    kern = am ? pairs_i8_kernel<ND, BN, BKB, ACC, STAGES, true, false, EW, COS, true, OutT, 2>
              : pairs_i8_kernel<ND, BN, BKB, ACC, STAGES, true, false, EW, COS, false, OutT, 2>;
  } else if constexpr (FAST) {
    kern = am ? pairs_i8_kernel<ND, BN, BKB, ACC, STAGES, true, false, EW, COS, true, OutT>
              : (sym ? pairs_i8_kernel<ND, BN, BKB, ACC, STAGES, true, true, EW, COS, false, OutT>
                     : pairs_i8_kernel<ND, BN, BKB, ACC, STAGES, true, false, EW, COS, false, OutT>);
  } else {
    kern = pairs_i8_kernel<ND, BN, BKB, ACC, STAGES, false, false, EW, COS, false, OutT>;
  }
  SOAP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
  const char *dbg = getenv("SOAP_I8_DEBUG");  // experiments: 1 = skip the epilogue, 2 = one MMA per K block, 4 = no stores
  {
    const long long excl = am ? am->excl : 0;
    const int dbg_flags = dbg ? atoi(dbg) : 0, kb = Dp / BKB;
    if constexpr (CG == 2) {
      // one cluster of two CTAs per TPC: as many pairs as the device can hold at once (a GPC with an odd number of
      // SMs leaves one SM idle), never more than there are tiles
      cudaLaunchConfig_t cfg;
      memset(&cfg, 0, sizeof(cfg));
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.gridDim = dim3(2 * (unsigned)(num_sms() / 2), 1, 1);
      cfg.blockDim = dim3(64 + 32 * EW, 1, 1);
      cfg.dynamicSmemBytes = SMEM;
      cfg.stream = s;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      int clusters = 0;
      SOAP_CUDA(cudaOccupancyMaxActiveClusters(&clusters, kern, &cfg));
      SOAP_REQUIRE(clusters >= 1, "soap_pairs: no CTA pair fits on this device");
      if ((long long)clusters > tiles) clusters = (int)tiles;
      grid = 2 * clusters;
      cfg.gridDim = dim3((unsigned)grid, 1, 1);
      ProfScope prof(SOAP_PROF_PAIRS, s);
      SOAP_CUDA(cudaLaunchKernelEx(&cfg, kern, map_a, map_b, out, (const double *)sa, (const double *)sb, (const double *)ua,
                                   (const double *)ub, (long long)M, (long long)N, kb, (long long)ldo, kind, power, tiles_m,
                                   tiles_n, dbg_flags, am_val, am_idx, excl));
      rc = check_launch("pairs_i8_kernel (CTA pairs)");
    } else {
      ProfScope prof(SOAP_PROF_PAIRS, s);
      kern<<<grid, 64 + 32 * EW, SMEM, s>>>(map_a, map_b, out, sa, sb, ua, ub, M, N, kb, ldo, kind, power, tiles_m, tiles_n,
                                          dbg_flags, am_val, am_idx, excl);
      rc = check_launch("pairs_i8_kernel");
    }
  }
  if (rc || !am) return rc;
  rowmin_reduce_kernel<<<(unsigned)((M + 255) / 256), 256, 0, s>>>(am_val, am_idx, (long long)tiles_n * PARTS, M, am->argmin_out,
                                                                 am->min_out);
  return check_launch("rowmin_reduce_kernel");
}

size_t pairs_i8_scratch(int64_t M, int64_t N, int D, int dtype) {
  // planes are padded to a multiple of 128 columns: valid for every K-block width used below
  return dtype == SOAP_F64 ? i8_scratch_bytes<6, 128>(M, N, D) : i8_scratch_bytes<3, 128>(M, N, D);
}

size_t pairs_i8_rowmin_scratch(int64_t M, int64_t N, int D, int dtype) {
  return pairs_i8_scratch(M, N, D, dtype) + i8_rowmin_bytes(M, N);
}

#define I8_LAUNCH(IN, ND, BN, BKB, ACC, ST, FAST, EW, COS, OUT) \
  launch_pairs_i8<IN, ND, BN, BKB, ACC, ST, FAST, EW, COS, OUT>(A, B, out, M, N, D, ldo, kind, power, scratch, s, am)
#define I8_LAUNCH_PAIR(IN, ND, BN, BKB, ACC, ST, EW, COS, OUT) \
  launch_pairs_i8<IN, ND, BN, BKB, ACC, ST, true, EW, COS, OUT, 2>(A, B, out, M, N, D, ldo, kind, power, scratch, s, am)

// CTA pairs (cta_group::2) are OPT-IN (SOAP_I8_CG=2): bit-identical to single CTAs on every shape tried
// (scripts/pairs_cg_check.py), but not faster -- 25 000 x 200 000 x 576 block: fp32 15.1 against 14.5 ms, fp64 65.4
// against 63.8; MMAs alone (25 000^2): fp64 5.32 against 5.41 ms, fp32 1.31 against 1.22.  The N = 80 passes are NOT
// shared-memory-operand bound as round 1 assumed: halving the B reads per CTA changes nothing, the int8 pipe itself
// delivers ~5000-6300 MACs/clk/SM here (cuBLASLt IMMA measured in the bench: 5400), not the nominal 8192.  Pairs
// couple the epilogues of two SMs (the next MMA waits for both), which costs more than the smaller stages give.
// The symmetric call (data is spectra) keeps its own single-CTA tile walk.
static bool i8_use_pairs(const void *A, const void *B, int64_t M, int64_t N, const I8RowMin *am) {
  const char *cg = getenv("SOAP_I8_CG");
  if (!cg || atoi(cg) != 2) return false;
  const bool same = (A == B && M == N);
  const bool sym = same && am == nullptr && getenv("SOAP_I8_NO_SYMMETRY") == nullptr;
  return !sym && M > kI8BM;
}

// Tile shapes measured on 25 000^2 x 576 (ms, first version of the epilogue), fp32: N160/K64/4 stages 2.19 |
// N80/K64/4 stages, 2 accumulator sets 2.33 | N160/K128/2 stages 2.35 | N80/K128/2 stages, 2 sets 2.41;
// fp64: N80/K64/2 stages is the best of N32 (2 sets) / N80 K32 4 stages / N64 3 stages.
int pairs_i8_f32(const float *A, const float *B, float *out, int64_t M, int64_t N, int D, int64_t ldo, int kind,
                 int power, void *scratch, cudaStream_t s, int32_t *argmin_out, double *min_out, long long excl) {
  I8RowMin rm{argmin_out, min_out, excl};
  const I8RowMin *am = (argmin_out || min_out) ? &rm : nullptr;
  if (kind == SOAP_KIND_SIMPLE || kind == SOAP_KIND_KERNEL) return I8_LAUNCH(float, 3, 160, 64, 1, 3, true, 8, true, float);
  if (kind == SOAP_KIND_NORMALIZED) {
    // SOAP_I8_TILE=80: two accumulator sets of N = 80 (the epilogue of a tile overlaps the MMAs of the next one; the
    // N = 80 MMAs are shared-memory-operand bound) instead of one set of N = 160 (experiments, scripts/prof_block.py)
    const char *tile = getenv("SOAP_I8_TILE");
    if (tile && atoi(tile) == 80) return I8_LAUNCH(float, 3, 80, 64, 2, 4, true, 8, false, float);
    const char *stg = getenv("SOAP_I8_STAGES");  // experiments: smaller operand pipelines
    if (stg && atoi(stg) == 2) return I8_LAUNCH(float, 3, 160, 64, 1, 2, true, 8, false, float);
    if (stg && atoi(stg) == 432) return I8_LAUNCH(float, 3, 160, 32, 1, 4, true, 8, false, float);
    if (i8_use_pairs(A, B, M, N, am)) return I8_LAUNCH_PAIR(float, 3, 160, 64, 1, 4, 8, false, float);
    return I8_LAUNCH(float, 3, 160, 64, 1, 3, true, 8, false, float);
  }
  return I8_LAUNCH(float, 3, 80, 128, 2, 2, false, 8, true, float);
}

int pairs_i8_f64(const double *A, const double *B, double *out, int64_t M, int64_t N, int D, int64_t ldo, int kind,
                 int power, void *scratch, cudaStream_t s, int32_t *argmin_out, double *min_out, long long excl) {
  I8RowMin rm{argmin_out, min_out, excl};
  const I8RowMin *am = (argmin_out || min_out) ? &rm : nullptr;
  if (kind == SOAP_KIND_SIMPLE || kind == SOAP_KIND_KERNEL) return I8_LAUNCH(double, 6, 80, 64, 1, 2, true, 12, true, double);
  if (kind == SOAP_KIND_NORMALIZED) {
    if (i8_use_pairs(A, B, M, N, am)) return I8_LAUNCH_PAIR(double, 6, 80, 64, 1, 3, 12, false, double);
    return I8_LAUNCH(double, 6, 80, 64, 1, 2, true, 12, false, double);
  }
  return I8_LAUNCH(double, 6, 80, 64, 1, 2, false, 8, true, double);
}
#undef I8_LAUNCH
#undef I8_LAUNCH_PAIR

}  // namespace soapb
