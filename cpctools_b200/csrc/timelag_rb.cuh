// timelag_rb.cuh -- the REGISTER-BLOCKED time-lag sweep (included by timelag.cu).
//
// Same arithmetic and outputs as timelag_kernel (sums[a][q][f] = sum_e m_e x[f][a][e] x[f-w_q][a][e], q = 0 is
// the squared norm), reference semantics timeSOAP, src/SOAPify/analysis.py:55-60.  What changes is who reads what
// from shared memory.  In timelag_kernel a lane owns ONE frame and reads its own unit plus one unit per window:
// 1 + NW shared-memory words per frame and unit, and with four windows the shared-memory pipe (128 B/clk), not
// HBM, sets the pace (ncu: 84 % busy at 0.71 of the HBM peak).  Here a lane owns kRbJ = 7 CONSECUTIVE frames of
// one 16-byte unit and keeps them in registers, together with the 10 frames before them: every window up to 10
// frames finds its partner in a register, a longer window costs one word per frame.  For windows (1,5,10,50):
// 7 + 10 + 7 = 24 words per 7 frames instead of 35.
//
// Geometry.  A half warp = 16 groups of 7 frames = one BLOCK of 112 frames; the 8 half warps of the 4 consumer
// warps take the 8 units of a 128-byte LINE, lane (half warp hw, group g) reading unit (hw + g) & 7 so that the 8
// lanes of a quarter warp hit 8 different bank groups although their rows are multiples of 128 bytes apart.
// Pieces are whole LINES of the frame row, aligned to the 128-byte grid of the tensor (an atom's row may start
// anywhere in its first line: the multiplicity vector is shifted by that phase and zero outside the row), so a
// tensor-map box never fetches a partial line: DRAM traffic = the algorithmic bytes (the unaligned 1648-byte
// pieces of timelag_kernel fetch 7.5 % more; scripts/micro/tma_stream.cu: 1.06 vs 0.99 of the copy peak).
//
// Ring.  Per stage (a group of lines of the piece) kRbR = 168 frame slots: the block being consumed (112) + the
// 56 frames before it (windows up to 56).  The next block's frames overwrite slots the current block still
// needs, so they are fetched STAGE BY STAGE: the box of (block b+1, stage s) is issued as soon as all consumer
// warps are done with (block b, stage s) -- "column pipelining": the ring needs no second block buffer, which is
// what leaves room for pieces of 6..8 lines (with a double-buffered block a piece would be 3 lines and the
// per-block hand-over of the partial sums would cost more than the saved loads).
#pragma once

namespace soapb {

constexpr int kRbJ = 7;                 // frames per lane
constexpr int kRbG = 16;                // groups per half warp
constexpr int kRbB = kRbJ * kRbG;       // 112 frames per block
constexpr int kRbBox = 56;              // frames per tensor-map box (two per block, three per ring)
constexpr int kRbR = 3 * kRbBox;        // ring slots per stage
constexpr int kRbCW = 4;                // consumer warps (8 half warps = the 8 units of a line)
constexpr int kRbSets = 2 * kRbCW;      // partial-sum sets per block (one per half warp)
constexpr int kRbHist = 10;             // windows up to this many frames are served from registers
constexpr int kRbMaxLag = kRbBox;       // longest window the ring keeps
constexpr int kRbMaxStages = 4;
constexpr int kRbThreads = (kRbCW + 2) * 32;

struct RbStages {
  int n;                          // stages per piece
  int k[kRbMaxStages];            // lines per stage
  int line0[kRbMaxStages];        // first line of the stage within the piece
  unsigned region[kRbMaxStages];  // byte offset of the stage's ring in shared memory
};

struct RbPlan {
  int n_units, c, pieces, Q;
  RbStages st;
  size_t ring_bytes, smem;
  size_t mult_elems;  // elements of one phase of the shifted multiplicity table
};

// lines of a piece: as few padded lines as possible over the 8 possible phases, then the widest piece
static bool rb_make_plan(int D, int elem_size, int nw, RbPlan &pl) {
  const int ue = 16 / elem_size;
  if (D % ue) return false;
  pl.n_units = D / ue;
  pl.Q = nw + 1;
  const int lines = (7 + pl.n_units + 7) / 8;  // worst phase
  const size_t budget = (size_t)max_smem_optin() - 1024;
  const size_t fixed = (size_t)kRbSets * pl.Q * kRbB * 8 + 256;
  const int force_c = env_int("SOAP_RB_LINES", 0), force_st = env_int("SOAP_RB_STAGES", 0);
  int best_c = 0;
  double best_waste = 1e30;
  for (int c = 2; c <= 16; ++c) {
    if (force_c > 0 && c != force_c) continue;
    if ((size_t)kRbR * c * 128 + (size_t)c * 128 + fixed > budget) continue;
    const int pieces = (lines + c - 1) / c;
    const double waste = (double)(pieces * c - lines) / lines + 0.004 * pieces;  // padded lines, then hand-overs
    if (waste < best_waste - 1e-12 || (waste < best_waste + 1e-12 && c > best_c)) best_waste = waste, best_c = c;
  }
  if (best_c == 0) return false;
  pl.c = best_c;
  pl.pieces = (lines + best_c - 1) / best_c;
  int ns = force_st > 0 ? force_st : (best_c >= 5 ? 3 : 2);
  if (ns > best_c) ns = best_c;
  if (ns > kRbMaxStages) ns = kRbMaxStages;
  if ((best_c + ns - 1) / ns > 4) ns = (best_c + 3) / 4;  // a box is at most 4 lines wide
  if (ns > kRbMaxStages) return false;
  pl.st.n = ns;
  unsigned off = 0;
  int line = 0;
  for (int s = 0; s < kRbMaxStages; ++s) {
    const int k = s < ns ? (best_c - line + (ns - s) - 1) / (ns - s) : 0;  // as even as possible, larger first
    pl.st.k[s] = k;
    pl.st.line0[s] = line;
    pl.st.region[s] = off;
    off += (unsigned)kRbR * k * 128;
    line += k;
  }
  pl.ring_bytes = off;
  pl.smem = pl.ring_bytes + (size_t)best_c * 128 + fixed;
  pl.mult_elems = (size_t)pl.pieces * best_c * 8 * ue;
  return true;
}

// shifted multiplicities: table[ph][u'] = mult[u' - ph] for units inside the row, 0 outside (ph = 0..7)
template <typename T>
__global__ void rb_mult_setup_kernel(const T *__restrict__ mult, T *__restrict__ table, int D, int units_total) {
  constexpr int UE = 16 / (int)sizeof(T);
  const int n = 8 * units_total * UE;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int ph = i / (units_total * UE), r = i - ph * units_total * UE;
    const int e = r - ph * UE;
    table[i] = (e >= 0 && e < D) ? (mult ? mult[e] : (T)1) : (T)0;
  }
}

template <typename T> struct RbVec;
template <> struct RbVec<double> {
  typedef double2 type;
  static __device__ __forceinline__ type lds(uint32_t a) { return lds_d2(a); }
  static __device__ __forceinline__ type zero() { return make_double2(0.0, 0.0); }
};
template <> struct RbVec<float> {
  typedef float4 type;
  static __device__ __forceinline__ type lds(uint32_t a) { return lds_f4(a); }
  static __device__ __forceinline__ type zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
};

// acc += sum_e xm_e * p_e over the elements of one unit.  fp64: exact products, DFMA.  fp32 data: the 4
// products of the unit are chained in fp32 and the chain is promoted to double (shorter chains than the 8-term
// ones of timelag_kernel: same or smaller rounding)
__device__ __forceinline__ void rb_dot(double &acc, const double2 &xm, const double2 &p) {
  acc = fma(xm.y, p.y, fma(xm.x, p.x, acc));
}
__device__ __forceinline__ void rb_dot(double &acc, const float4 &xm, const float4 &p) {
  acc += (double)fmaf(xm.w, p.w, fmaf(xm.z, p.z, fmaf(xm.y, p.y, xm.x * p.x)));
}
__device__ __forceinline__ double2 rb_mul(const double2 &x, const double2 &m) { return make_double2(x.x * m.x, x.y * m.y); }
__device__ __forceinline__ float4 rb_mul(const float4 &x, const float4 &m) {
  return make_float4(x.x * m.x, x.y * m.y, x.z * m.z, x.w * m.w);
}

// units outside the atom's row (first / last line of a row that does not start or end on a line) carry
// multiplicity 0; their DATA belongs to a neighbouring atom and must not leak even as 0 * NaN
__device__ __forceinline__ void rb_clean(double2 &v, const double2 &m) {
  v.x = m.x != 0.0 ? v.x : 0.0;
  v.y = m.y != 0.0 ? v.y : 0.0;
}
__device__ __forceinline__ void rb_clean(float4 &v, const float4 &m) {
  v.x = m.x != 0.f ? v.x : 0.f;
  v.y = m.y != 0.f ? v.y : 0.f;
  v.z = m.z != 0.f ? v.z : 0.f;
  v.w = m.w != 0.f ? v.w : 0.f;
}

// window W <= kRbHist: the partner of frame i is frame i - W of the same lane or one of the 10 frames before
template <int W, typename V>
__device__ __forceinline__ void rb_small_lag(double (&acc)[kRbJ], const V (&xm)[kRbJ], const V (&x)[kRbJ],
                                             const V (&h)[kRbHist]) {
#pragma unroll
  for (int i = 0; i < kRbJ; ++i) rb_dot(acc[i], xm[i], i >= W ? x[i >= W ? i - W : 0] : h[i >= W ? 0 : W - 1 - i]);
}
template <typename V>
__device__ __forceinline__ void rb_small_lag_dispatch(int w, double (&acc)[kRbJ], const V (&xm)[kRbJ], const V (&x)[kRbJ],
                                                      const V (&h)[kRbHist]) {
  switch (w) {  // w is uniform over the grid: one branch per window and unit
    case 1: rb_small_lag<1>(acc, xm, x, h); break;
    case 2: rb_small_lag<2>(acc, xm, x, h); break;
    case 3: rb_small_lag<3>(acc, xm, x, h); break;
    case 4: rb_small_lag<4>(acc, xm, x, h); break;
    case 5: rb_small_lag<5>(acc, xm, x, h); break;
    case 6: rb_small_lag<6>(acc, xm, x, h); break;
    case 7: rb_small_lag<7>(acc, xm, x, h); break;
    case 8: rb_small_lag<8>(acc, xm, x, h); break;
    case 9: rb_small_lag<9>(acc, xm, x, h); break;
    default: rb_small_lag<10>(acc, xm, x, h); break;
  }
}

// Warp roles: warps 0..3 consume, warp 4 produces (TMA), warp 5 reduces the 8 partial-sum sets of a block and adds
// up the pieces of an atom in the per-CTA accumulator (L2-resident, as in timelag_kernel).  mbarriers only:
//   full[s]     producer -> consumers   stage s of the current block has landed (+ the item's multiplicities)
//   empty[s]    consumers -> producer   stage s of the current block consumed: its slots may take the next block
//   red_full / red_empty                partial sums of a block handed to the reducer / taken
// Windows are sorted: win[0..NS) <= kRbHist (registers), win[NS..NS+NL) <= kRbMaxLag (one load per frame).
template <typename T, int NS, int NL>
__global__ void __launch_bounds__(kRbThreads, 1)
    timelag_rb_kernel(const T *__restrict__ mult_table, double *__restrict__ sums, double *__restrict__ cta_acc, long long F,
                      long long A, int n_units, int c, int pieces, int mult_elems, int4 win4, int hist_depth,
                      int diag, const RbStages st, const __grid_constant__ CUtensorMap tm1, const __grid_constant__ CUtensorMap tm2,
                      const __grid_constant__ CUtensorMap tm3, const __grid_constant__ CUtensorMap tm4) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  typedef typename RbVec<T>::type V;
  constexpr int NW = NS + NL, Q = NW + 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int win[4] = {win4.x, win4.y, win4.z, win4.w};

  unsigned char *ring = smem_raw;
  unsigned ring_bytes = 0;
#pragma unroll
  for (int s = 0; s < kRbMaxStages; ++s) ring_bytes += (unsigned)kRbR * st.k[s] * 128;
  unsigned char *msm = ring + ring_bytes;                                 // [c lines][128 B] multiplicities of the item
  double *red = reinterpret_cast<double *>(msm + (size_t)c * 128);        // [kRbSets][Q][kRbB]
  uint64_t *full = reinterpret_cast<uint64_t *>(red + kRbSets * Q * kRbB);  // [kRbMaxStages]
  uint64_t *empty = full + kRbMaxStages;                                  // [kRbMaxStages]
  uint64_t *red_full = empty + kRbMaxStages;
  uint64_t *red_empty = red_full + 1;

  const long long my_atoms = (long long)blockIdx.x < A ? (A - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int n_my = (int)my_atoms * pieces;
  const int nblk = (int)((F + kRbB - 1) / kRbB);

  if (tid == 0) {
    for (int s = 0; s < kRbMaxStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kRbCW);
    }
    mbar_init(red_full, kRbCW);
    mbar_init(red_empty, 1);
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == kRbCW) {
    // ===== producer ==========================================================================
    if (lane == 0) {
      const CUtensorMap *maps[4] = {&tm1, &tm2, &tm3, &tm4};
      int gb = 0;
      for (int k = 0; k < n_my; ++k) {
        const long long a = (long long)blockIdx.x + (long long)(k / pieces) * gridDim.x;
        const int p = k % pieces;
        const long long u0 = a * n_units;            // first unit of the atom's row in the frame row
        const long long line_a = u0 >> 3;            // its first line
        const int ph = (int)(u0 & 7);                // where in that line the row starts
        const T *mrow = mult_table + (size_t)ph * mult_elems;
        for (int b = 0; b < nblk; ++b, ++gb) {
          const int slot0 = (gb * kRbB) % kRbR;      // 0, 112, 56, 0, ...: boxes of 56 slots never wrap
          const int slot1 = slot0 + kRbBox >= kRbR ? slot0 + kRbBox - kRbR : slot0 + kRbBox;
          for (int s = 0; s < st.n; ++s) {
            if (gb >= 1) mbar_wait(&empty[s], (uint32_t)((gb - 1) & 1));
            const uint32_t row_bytes = (uint32_t)st.k[s] * 128u;
            const uint32_t mbytes = b == 0 ? row_bytes : 0u;
            if (diag & 2) {  // diagnostic: no copies
              mbar_arrive(&full[s]);
              continue;
            }
            mbar_expect_tx(&full[s], 2u * kRbBox * row_bytes + mbytes);
            const int w0 = (int)((line_a + (long long)p * c + st.line0[s]) * 16);
            unsigned char *reg = ring + st.region[s];
            tma_load_2d(reg + (size_t)slot0 * row_bytes, maps[st.k[s] - 1], w0, b * kRbB, &full[s]);
            tma_load_2d(reg + (size_t)slot1 * row_bytes, maps[st.k[s] - 1], w0, b * kRbB + kRbBox, &full[s]);
            if (mbytes)
              bulk_g2s(msm + (size_t)st.line0[s] * 128, reinterpret_cast<const unsigned char *>(mrow) + ((size_t)p * c + st.line0[s]) * 128,
                       mbytes, &full[s]);
          }
        }
      }
    }
  } else if (warp == kRbCW + 1) {
    // ===== reducer ===========================================================================
    int gb = 0;
    const uint64_t pol = l2_policy_evict_last();
    const size_t fpad = (size_t)nblk * kRbB;
    double *acc_row = cta_acc + (size_t)blockIdx.x * Q * fpad;
    for (int k = 0; k < n_my; ++k) {
      const long long a = (long long)blockIdx.x + (long long)(k / pieces) * gridDim.x;
      const int p = k % pieces;
      double *srow = sums + (size_t)a * Q * (size_t)F;
      const bool first_piece = p == 0, last_piece = p == pieces - 1;
      for (int b = 0; b < nblk; ++b, ++gb) {
        // a lane takes PAIRS of frames (2 * fl2, 2 * fl2 + 1), fl2 = lane, lane + 32 (< 56).  All loads of a pair
        // are issued before the sums (8 sets x Q = 40 independent 128-bit loads in flight; the first version
        // alternated load and dependent add, 160 round trips per block, and paced the whole CTA)
        double2 prev[2][Q];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int fl2 = lane + 32 * j;
#pragma unroll
          for (int q = 0; q < Q; ++q) {
            prev[j][q] = make_double2(0.0, 0.0);
            if (!first_piece && fl2 < kRbB / 2) {
              const double *ap = acc_row + (size_t)q * fpad + (size_t)b * kRbB + 2 * fl2;
              prev[j][q] = make_double2(ld_acc(ap, pol), ld_acc(ap + 1, pol));
            }
          }
        }
        mbar_wait(red_full, (uint32_t)(gb & 1));
        const uint32_t red_addr = smem_u32(red);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int fl2 = lane + 32 * j;
          if (fl2 < kRbB / 2) {
            double2 v[Q][kRbSets];
#pragma unroll
            for (int q = 0; q < Q; ++q)
#pragma unroll
              for (int w = 0; w < kRbSets; ++w) v[q][w] = lds_d2(red_addr + (uint32_t)(((w * Q + q) * kRbB + 2 * fl2) * 8));
            const long long f = (long long)b * kRbB + 2 * fl2;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
              double s0 = v[q][0].x, s1 = v[q][0].y;
#pragma unroll
              for (int w = 1; w < kRbSets; ++w) s0 += v[q][w].x, s1 += v[q][w].y;  // fixed order
              if (!first_piece) s0 = prev[j][q].x + s0, s1 = prev[j][q].y + s1;    // ((p_0 + p_1) + p_2) + ...
              if (last_piece) {
                if (f < F) srow[(size_t)q * F + f] = s0;
                if (f + 1 < F) srow[(size_t)q * F + f + 1] = s1;
              } else {
                st_acc(acc_row + (size_t)q * fpad + (size_t)f, s0, pol);
                st_acc(acc_row + (size_t)q * fpad + (size_t)f + 1, s1, pol);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(red_empty);
      }
    }
  } else {
    // ===== consumers: lane <-> 7 consecutive frames of one unit ===================================
    const int half = lane >> 4, g = lane & 15;
    const int hw = 2 * warp + half;
    const uint32_t unit_off = (uint32_t)((hw + g) & 7) * 16u;
    const uint32_t ring_addr = smem_u32(ring), msm_addr = smem_u32(msm) + unit_off;
    int gb = 0;
    for (int k = 0; k < n_my; ++k) {
      const long long u0 = ((long long)blockIdx.x + (long long)(k / pieces) * gridDim.x) * n_units;
      const int ph = (int)(u0 & 7), pline0 = (k % pieces) * c;
      const int first_partial = ph != 0 ? 0 : -1, last_whole = (ph + n_units) >> 3;  // lines >= last_whole are masked
      for (int b = 0; b < nblk; ++b, ++gb) {
        int s0 = (gb * kRbB) % kRbR + g * kRbJ;  // ring slot of this lane's first frame (a group never straddles
        if (s0 >= kRbR) s0 -= kRbR;              // the end of the ring: 168 = 24 groups)
        int sh[kRbHist], sy[NL > 0 ? NL : 1][kRbJ];
#pragma unroll
        for (int i = 0; i < kRbHist; ++i) sh[i] = s0 - 1 - i < 0 ? s0 - 1 - i + kRbR : s0 - 1 - i;
#pragma unroll
        for (int n = 0; n < NL; ++n)
#pragma unroll
          for (int i = 0; i < kRbJ; ++i) {
            const int s = s0 + i - win[NS + n];
            sy[n][i] = s < 0 ? s + kRbR : s;
          }
        double acc[Q][kRbJ];
#pragma unroll
        for (int q = 0; q < Q; ++q)
#pragma unroll
          for (int i = 0; i < kRbJ; ++i) acc[q][i] = 0.0;
        for (int s = 0; s < st.n; ++s) {
          const uint32_t stride = (uint32_t)st.k[s] * 128u;
          const uint32_t reg = ring_addr + st.region[s] + unit_off;
          const uint32_t mreg = msm_addr + (uint32_t)st.line0[s] * 128u;
          mbar_wait(&full[s], (uint32_t)(gb & 1));
          for (int l = 0; l < ((diag & 1) ? 0 : st.k[s]); ++l) {  // diag 1: no accumulation
            const uint32_t rb = reg + (uint32_t)l * 128u;
            V x[kRbJ], h[kRbHist], y[NL > 0 ? NL : 1][kRbJ];
            const V m = RbVec<T>::lds(mreg + (uint32_t)l * 128u);
#pragma unroll
            for (int i = 0; i < kRbJ; ++i) x[i] = RbVec<T>::lds(rb + (uint32_t)(s0 + i) * stride);
#pragma unroll
            for (int i = 0; i < kRbHist; ++i) h[i] = (NS > 0 && i < hist_depth) ? RbVec<T>::lds(rb + (uint32_t)sh[i] * stride) : RbVec<T>::zero();
#pragma unroll
            for (int n = 0; n < NL; ++n)
#pragma unroll
              for (int i = 0; i < kRbJ; ++i) y[n][i] = RbVec<T>::lds(rb + (uint32_t)sy[n][i] * stride);
            const int pl = pline0 + st.line0[s] + l;  // line of the atom's row
            if (pl == first_partial || pl >= last_whole) {
#pragma unroll
              for (int i = 0; i < kRbJ; ++i) rb_clean(x[i], m);
#pragma unroll
              for (int i = 0; i < kRbHist; ++i) rb_clean(h[i], m);
#pragma unroll
              for (int n = 0; n < NL; ++n)
#pragma unroll
                for (int i = 0; i < kRbJ; ++i) rb_clean(y[n][i], m);
            }
            V xm[kRbJ];
#pragma unroll
            for (int i = 0; i < kRbJ; ++i) {
              xm[i] = rb_mul(x[i], m);
              rb_dot(acc[0][i], xm[i], x[i]);
            }
#pragma unroll
            for (int n = 0; n < NS; ++n) rb_small_lag_dispatch<V>(win[n], acc[1 + n], xm, x, h);
#pragma unroll
            for (int n = 0; n < NL; ++n)
#pragma unroll
              for (int i = 0; i < kRbJ; ++i) rb_dot(acc[1 + NS + n][i], xm[i], y[n][i]);
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[s]);
        }
        // hand the partial sums of this half warp to the reducer
        if (gb >= 1) mbar_wait(red_empty, (uint32_t)((gb - 1) & 1));
        double *r = red + (size_t)hw * Q * kRbB + g * kRbJ;
#pragma unroll
        for (int q = 0; q < Q; ++q)
#pragma unroll
          for (int i = 0; i < kRbJ; ++i) r[q * kRbB + i] = acc[q][i];
        __syncwarp();
        if (lane == 0) mbar_arrive(red_full);
      }
    }
  }
}

}  // namespace soapb
