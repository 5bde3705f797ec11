// timelag_chain.cuh -- the time-lag sweep with LANE = UNITS OF THE ROW and the lag history in REGISTERS
// (included by timelag.cu).
//
// Same arithmetic and outputs as timelag_kernel (sums[a][q][f] = sum_e m_e x[f][a][e] x[f-w_q][a][e], q = 0 is the
// squared norm; reference semantics timeSOAP, src/SOAPify/analysis.py:55-60).  timelag_kernel (lane = frame) reads
// every staged element 1 + NW times from shared memory -- 26.55 shared-memory cycles per 512 B of HBM data for the
// four windows (1, 5, 10, 50), against ~20 cycles the HBM needs to deliver them: the sweep is bound by
// the SM's shared memory.  Here the windows that are multiples of 5 never touch shared memory:
//   * consumer warp r (of 5) owns the frames f = r (mod 5) of its item: a CHAIN r, r + 5, r + 10, ...;
//   * lane l owns the units l, l + 32, ... (U of them, 16 bytes each) of the piece and keeps m * x of the last TEN
//     chain frames of those units in registers: the partners for windows 5c (c = 1..10) are register operands;
//   * only the remaining window (1 in the headline sweep) is read from the ring, which therefore holds just the
//     landing blocks of the TMA engine (3 x 25 frames) instead of a 50-frame history: a piece is 128 units = 2 KB
//     (5 pieces per 9.8 kB row instead of 6, cut at multiples of 2 KB from the row start, the narrower last piece with
//     its own tensor map) and every 128-bit load of a warp is one contiguous 512-byte row segment (conflict free
//     whatever the pitch);
//   * nothing of a block is held back for the next one: the window-1 partner of a block's first frame is the LAST frame
//     of the block before, which chain 4's warp holds in registers and hands to chain 0 through a small shared-memory
//     halo (other ring windows <= 2: the boxes overlap by that many rows; 3..25: the previous buffer is kept).
// The price is a cross-lane sum: a lane ends a frame with Q partial sums over ITS units.  They go through shared memory
// (red[frame][sum][lane]) to two reducer warps, which add the 32 lanes in a fixed tree and then the pieces of the atom
// (bitwise reproducible across grid sizes and atom counts).  A shuffle butterfly inside the consumer warps (ch_reduce,
// SR = false: kept, not instantiated) was 2x slower: with 255 registers per thread only five consumer warps fit, one
// per scheduler, and a lone warp cannot hide five dependent shuffle + add levels.
// Per 512 B of HBM data: 4 (own frame) + 4 (window 1) + 2.5 + 2.5 (partial sums) + 4 (TMA write) = 17 shared-memory
// cycles against 26.55.  Measured: 0.96-0.99 of the HBM copy peak alone, 0.86-0.89 sustained under the
// power cap, against 0.78 / 0.72-0.75 for timelag_kernel.
// The history index must be static (registers cannot be indexed): the frame loop is unrolled over 10 chain steps
// = 2 blocks of 25 frames.
// Applies to: TMA-able rows of whole units, cosine-type kinds, atom counts that fill the SMs (a CTA owns whole atoms),
// windows = up to one window <= 25 that is no multiple of 5 + up to three distinct multiples of 5 up to 50, in one of
// the instantiated patterns (chain_choose).  Everything else takes timelag_kernel / timelag_pair_kernel.
#pragma once

namespace soapb {

constexpr int kChS = 5;      // chain stride in frames == consumer warps
constexpr int kChH = 10;     // chain steps of history in registers (windows up to 50)
constexpr int kChSteps = 5;  // chain steps per block
constexpr int kChB = kChS * kChSteps;  // frames per block (25)
constexpr int kChNbufMax = 4;  // ring blocks (as many as fit, at least 3: current + in flight, + the previous one for ring windows 3..25)
constexpr int kChRed = 2;     // reducer warps in the CTA (the launch decides how many of them work: nred)
constexpr int kChThreads = (kChS + 1 + kChRed) * 32;
constexpr int kChRedPitch = 34;  // doubles per (frame, sum) row of the lanes' partial sums (32 + 2: conflict-free 128-bit reads)

struct ChainPlan {
  int unit_elems, n_units, split, upp, pitch, last_units, Q, nbuf;
  size_t blk_bytes, smem;
};

// SR: the lanes' partial sums go through shared memory to the reducer warp (instead of the shuffle butterfly)
// ov: rows of the previous block fetched again in front of every box (the ring-window partners of a block's first
// frames then sit in the block's own buffer and nothing is held back for them: one more block in flight)
// halo: instead of fetching that row again, the warp of chain 4 hands the LAST frame of every block (which it holds in
// registers) to chain 0 through a small shared-memory area (4 x one piece): the boxes do not overlap (ov = 0) and
// nothing is held back either
constexpr int kChHalo = 4;  // halo buffers (the writer can be at most two blocks ahead of the reader)
static bool ch_make_plan(int D, int elem_size, int nw, int U, bool SR, int ov, bool halo, ChainPlan &pl) {
  pl.unit_elems = 16 / elem_size;
  if (D % pl.unit_elems) return false;
  pl.n_units = D / pl.unit_elems;
  pl.Q = 1 + nw;
  // pieces of 32 * U units (every lane busy, and 512 * U bytes apart: box rows start and end on 64-byte DRAM atoms
  // whenever the rows themselves do) + a narrower last piece with its own tensor map (dense box rows: its own pitch)
  pl.split = (pl.n_units + 32 * U - 1) / (32 * U);
  pl.upp = pl.split > 1 ? 32 * U : pl.n_units;
  pl.last_units = pl.n_units - (pl.split - 1) * pl.upp;
  pl.pitch = 16 * pl.upp;
  pl.blk_bytes = align_up((size_t)(kChB + ov) * pl.pitch, 128);  // TMA destinations are 128-byte aligned
  for (pl.nbuf = kChNbufMax; pl.nbuf >= 3; --pl.nbuf) {
    pl.smem = (size_t)pl.nbuf * pl.blk_bytes + (size_t)2 * kChB * pl.Q * 8 * (SR ? kChRedPitch : 1) + (halo ? (size_t)kChHalo * 512 * U : 0) +
              (size_t)(2 * kChNbufMax + 4 + kChHalo) * 8 + 64;
    if (pl.smem <= (size_t)max_smem_optin() - 1024) return true;
  }
  return false;
}

// sum over the 32 lanes of Q values each, transposing: the lanes end up holding DIFFERENT totals (halving the
// number of values a lane carries at every level instead of moving all Q through all 5 levels).  Returns the total
// this lane is left with in v[0]; ch_red_owner() tells which.  Fixed order: the result does not depend on anything
// but the 32 x Q inputs.
template <int N>
__device__ __forceinline__ void ch_reduce_level(double (&v)[N], int lane, int mask) {
  constexpr int KEEP = (N + 1) / 2;
  const bool hi = (lane & mask) != 0;
#pragma unroll
  for (int i = 0; i < KEEP; ++i) {
    const double a = v[i], b = (KEEP + i < N) ? v[KEEP + i] : 0.0;
    const double send = hi ? a : b, keep = hi ? b : a;
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
  }
}
template <int Q>
__device__ __forceinline__ double ch_reduce(double (&v)[Q], int lane) {
  static_assert(Q >= 2 && Q <= 5, "1..4 windows");
  // level 1 (lanes 16 apart): Q -> ceil(Q/2) values per lane; level 2: -> ceil(ceil(Q/2)/2) <= 2; level 3: -> 1
  ch_reduce_level<Q>(v, lane, 16);
  constexpr int N2 = (Q + 1) / 2;
  double v2[N2];
#pragma unroll
  for (int i = 0; i < N2; ++i) v2[i] = v[i];
  if constexpr (N2 > 1) {
    ch_reduce_level<N2>(v2, lane, 8);
    constexpr int N3 = (N2 + 1) / 2;
    double v3[N3];
#pragma unroll
    for (int i = 0; i < N3; ++i) v3[i] = v2[i];
    if constexpr (N3 > 1) {
      ch_reduce_level<N3>(v3, lane, 4);
    } else {
      v3[0] += __shfl_xor_sync(0xffffffffu, v3[0], 4);
    }
    v2[0] = v3[0];
  } else {
    v2[0] += __shfl_xor_sync(0xffffffffu, v2[0], 8);
    v2[0] += __shfl_xor_sync(0xffffffffu, v2[0], 4);
  }
  double r = v2[0];
  r += __shfl_xor_sync(0xffffffffu, r, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}
// the sum index q whose total lane `lane` holds after ch_reduce (-1: none).  Level 1 splits q < K1 | q >= K1 by lane
// bit 16, level 2 by bit 8, level 3 by bit 4.
template <int Q>
__device__ __forceinline__ int ch_red_owner(int lane) {
  constexpr int K1 = (Q + 1) / 2;
  const int b16 = (lane >> 4) & 1, b8 = (lane >> 3) & 1, b4 = (lane >> 2) & 1;
  int lo = b16 ? K1 : 0, n = b16 ? Q - K1 : K1;  // this lane's range after level 1: [lo, lo + n)
  if constexpr (K1 > 1) {
    constexpr int K2 = (K1 + 1) / 2;
    // every lane runs the level with KEEP = K2 (the hi half of level 1 may carry fewer values: zeros fill in)
    if (b8) lo += K2, n -= K2; else n = n < K2 ? n : K2;
    if constexpr (K2 > 1) {
      constexpr int K3 = (K2 + 1) / 2;
      if (b4) lo += K3, n -= K3; else n = n < K3 ? n : K3;
    }
  }
  return n >= 1 ? lo : -1;
}

// ONE chain step of a lane: the U units of frame f (x_addr) against the window-w0 partner from the ring (p_addr) and
// the history h (m * x of the previous chain frames), then h[N] <- m * x.  v[q]: q = 0 squared norm, 1 = ring window
// (when NS), then the chain windows C1, C2, C3 (chain steps back; 0 = unused).
template <typename T> struct ChVec;
template <> struct ChVec<double> { using type = double2; };
template <> struct ChVec<float> { using type = float4; };

template <int N, int U, int NS, int C1, int C2, int C3, int Q>
__device__ __forceinline__ void ch_step_f64(uint32_t x_addr, uint32_t p_addr, const uint32_t col, const uint32_t (&cofs)[U],
                                            const double2 (&m)[U], double2 (&h)[kChH][U], double (&v)[Q], const uint32_t halo_st = 0) {
  // no predicates: a lane without a unit in column c reads the LAST unit of the piece instead (cofs) and multiplies it
  // by m = 0 -- data of the same atom and frame, so a NaN there poisons nothing that is not NaN already
  double2 x[U], p[U];
#pragma unroll
  for (int c = 0; c < U; ++c) {
    x[c] = lds_d2(x_addr + cofs[c]);
    if (NS) p[c] = lds_d2(p_addr + cofs[c]);
  }
  if (halo_st) {  // the last frame of the block, for the first frame of the next one (chain 0)
#pragma unroll
    for (int c = 0; c < U; ++c)
      asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(halo_st + col + 512u * c), "d"(x[c].x), "d"(x[c].y) : "memory");
  }
  // ONE chain per sum: the Q sums are Q independent chains already (a second chain per sum costs 10 registers and 5
  // additions per step)
#pragma unroll
  for (int q = 0; q < Q; ++q) v[q] = 0.0;
#pragma unroll
  for (int c = 0; c < U; ++c) {
    // the history terms first (they need neither m * x nor the ring partner, and the oldest slot is read before it is
    // overwritten), then m * x and the two sums that use it
    int q = 1 + NS;
    if (C1) {
      const double2 y = h[(N + kChH - C1) % kChH][c];
      v[q] = fma(x[c].y, y.y, fma(x[c].x, y.x, v[q]));
      ++q;
    }
    if (C2) {
      const double2 y = h[(N + kChH - C2) % kChH][c];
      v[q] = fma(x[c].y, y.y, fma(x[c].x, y.x, v[q]));
      ++q;
    }
    if (C3) {
      const double2 y = h[(N + kChH - C3) % kChH][c];
      v[q] = fma(x[c].y, y.y, fma(x[c].x, y.x, v[q]));
      ++q;
    }
    const double2 xm = make_double2(x[c].x * m[c].x, x[c].y * m[c].y);
    h[N][c] = xm;
    v[0] = fma(xm.y, x[c].y, fma(xm.x, x[c].x, v[0]));
    if (NS) v[1] = fma(xm.y, p[c].y, fma(xm.x, p[c].x, v[1]));
  }
}

// fp32 data: products and two 2U-term chains per sum in fp32, promoted to double before the cross-lane sum (the same
// accuracy class as accum_units_f32: the fp32 rounding is bounded to one short chain)
template <int N, int U, int NS, int C1, int C2, int C3, int Q>
__device__ __forceinline__ void ch_step_f32(uint32_t x_addr, uint32_t p_addr, const uint32_t col, const bool (&act)[U],
                                            const float4 (&m)[U], float4 (&h)[kChH][U], double (&v)[Q], const uint32_t halo_st = 0) {
  float4 x[U], p[U];
#pragma unroll
  for (int c = 0; c < U; ++c) {
    x[c] = act[c] ? lds_f4(x_addr + col + 512u * c) : make_float4(0.f, 0.f, 0.f, 0.f);
    if (NS) p[c] = act[c] ? lds_f4(p_addr + col + 512u * c) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (halo_st) {
#pragma unroll
    for (int c = 0; c < U; ++c)
      asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(halo_st + col + 512u * c), "f"(x[c].x), "f"(x[c].y), "f"(x[c].z), "f"(x[c].w) : "memory");
  }
  float s[Q][2];
#pragma unroll
  for (int q = 0; q < Q; ++q) s[q][0] = s[q][1] = 0.f;
  auto dot = [](const float4 &a, const float4 &b, float (&acc)[2]) {
    acc[0] = fmaf(a.x, b.x, acc[0]);
    acc[1] = fmaf(a.y, b.y, acc[1]);
    acc[0] = fmaf(a.z, b.z, acc[0]);
    acc[1] = fmaf(a.w, b.w, acc[1]);
  };
#pragma unroll
  for (int c = 0; c < U; ++c) {
    const float4 xm = make_float4(x[c].x * m[c].x, x[c].y * m[c].y, x[c].z * m[c].z, x[c].w * m[c].w);
    dot(xm, x[c], s[0]);
    int q = 1;
    if (NS) dot(xm, p[c], s[q++]);
    if (C1) dot(x[c], h[(N + kChH - C1) % kChH][c], s[q++]);
    if (C2) dot(x[c], h[(N + kChH - C2) % kChH][c], s[q++]);
    if (C3) dot(x[c], h[(N + kChH - C3) % kChH][c], s[q++]);
    h[N][c] = xm;
  }
#pragma unroll
  for (int q = 0; q < Q; ++q) v[q] = (double)s[q][0] + (double)s[q][1];
}

// ONE column (unit) of a chain step, for the software-pipelined block: the loads of column t + PD are issued before
// the arithmetic of column t, across the steps of a block (a lone warp per scheduler cannot hide a load -> use chain
// otherwise: measured 2x slower)
template <int N, int C, int U, int NS, int C1, int C2, int C3, int Q>
__device__ __forceinline__ void ch_col(const double2 &x, const double2 &p, const double2 &m, double2 (&h)[kChH][U], double (&s)[Q][2]) {
  const double2 xm = make_double2(x.x * m.x, x.y * m.y);
  s[0][0] = fma(xm.x, x.x, s[0][0]);
  s[0][1] = fma(xm.y, x.y, s[0][1]);
  int q = 1;
  if (NS) {
    s[q][0] = fma(xm.x, p.x, s[q][0]);
    s[q][1] = fma(xm.y, p.y, s[q][1]);
    ++q;
  }
  if (C1) {
    const double2 y = h[(N + kChH - C1) % kChH][C];
    s[q][0] = fma(x.x, y.x, s[q][0]);
    s[q][1] = fma(x.y, y.y, s[q][1]);
    ++q;
  }
  if (C2) {
    const double2 y = h[(N + kChH - C2) % kChH][C];
    s[q][0] = fma(x.x, y.x, s[q][0]);
    s[q][1] = fma(x.y, y.y, s[q][1]);
    ++q;
  }
  if (C3) {
    const double2 y = h[(N + kChH - C3) % kChH][C];
    s[q][0] = fma(x.x, y.x, s[q][0]);
    s[q][1] = fma(x.y, y.y, s[q][1]);
    ++q;
  }
  h[N][C] = xm;
}
template <int N, int C, int U, int NS, int C1, int C2, int C3, int Q>
__device__ __forceinline__ void ch_col(const float4 &x, const float4 &p, const float4 &m, float4 (&h)[kChH][U], float (&s)[Q][2]) {
  // (packed FFMA2 / FMUL2 need aligned register pairs: at the 255-register cap they spill, measured 0.46-0.48 of the
  // HBM peak against 0.61-0.68 unpacked)
  auto dot = [](const float4 &a, const float4 &b, float (&acc)[2]) {
    acc[0] = fmaf(a.x, b.x, acc[0]);
    acc[1] = fmaf(a.y, b.y, acc[1]);
    acc[0] = fmaf(a.z, b.z, acc[0]);
    acc[1] = fmaf(a.w, b.w, acc[1]);
  };
  const float4 xm = make_float4(x.x * m.x, x.y * m.y, x.z * m.z, x.w * m.w);
  dot(xm, x, s[0]);
  int q = 1;
  if (NS) dot(xm, p, s[q++]);
  if (C1) dot(x, h[(N + kChH - C1) % kChH][C], s[q++]);
  if (C2) dot(x, h[(N + kChH - C2) % kChH][C], s[q++]);
  if (C3) dot(x, h[(N + kChH - C3) % kChH][C], s[q++]);
  h[N][C] = xm;
}
__device__ __forceinline__ void ch_lds(double2 &v, uint32_t addr, bool on) { v = on ? lds_d2(addr) : make_double2(0.0, 0.0); }
__device__ __forceinline__ void ch_lds(float4 &v, uint32_t addr, bool on) { v = on ? lds_f4(addr) : make_float4(0.f, 0.f, 0.f, 0.f); }

template <typename F, int... I>
__device__ __forceinline__ void ch_static_for(F &&f, std::integer_sequence<int, I...>) {
  (f(std::integral_constant<int, I>{}), ...);
}

// Item of the chain kernel.  `align` (rows that start 64 bytes into a 128-byte line for every other atom: n_units % 8
// == 4): the pieces of such an atom are cut at p * upp - 4 units, so that every piece boundary inside a row falls on a
// 128-byte line; its first piece is 4 units shorter, its last one 4 units longer.  kind: 0 full piece, 1 the short
// first piece, 2 the last piece, 3 the long last piece (tensor map / pitch of the box).
struct ChItem {
  TlItem it;
  int kind;
};
template <int UE>
__device__ __forceinline__ ChItem ch_item(int k, int split, int upp, int D, int align) {
  ChItem ci;
  ci.it = tl_item<UE>(k, split, upp, D, true, false);
  const bool last = ci.it.p == split - 1;
  ci.kind = last ? 2 : 0;
  if (align) {
    const int n_units = D / UE;
    const int s4 = (int)((ci.it.a * n_units) & 7);  // 0 or 4
    if (s4) {
      const int u0 = ci.it.p == 0 ? 0 : ci.it.p * upp - s4;
      const int u1 = last ? n_units : (ci.it.p + 1) * upp - s4;
      constexpr int ES = 16 / UE;
      ci.it.e0 = u0 * UE;
      ci.it.punits = u1 - u0;
      ci.it.pe = ci.it.punits * UE;
      ci.it.w0 = (ci.it.a * D + ci.it.e0) * ES / 8;
      ci.kind = last ? 3 : (ci.it.p == 0 ? 1 : 0);
    }
  }
  return ci;
}

// Warp roles: warps 0..4 consume (warp = frame mod 5, lane = units), warp 5 produces (one tensor-map box of 25 frames
// x one piece per block), warps 6 and 7 add the lanes' partial sums and the pieces of an atom (L2-resident per-CTA
// accumulator, as timelag_kernel); with one reducer warp it was 81 % busy and the consumers waited for it 10 % of
// their time (profiles/r03_timelag_chain_ncu_full.txt), with two the sweep is ~4 % faster once the GPU is power capped
// (SOAP_TL_CHAIN_RED=1 | 2, interleaved A/B in profiles/r03_timelag_chain_ab2.txt).
//   full[s]        producer -> consumers   block in ring buffer s has landed
//   empty[s]       consumers -> producer   buffer s is no longer needed (arrived at the END of the following block:
//                                          its first frames read their ring-window partners from buffer s)
//   red_full[i]    consumers -> reducer    the sums of a block are in red[i] (i = block parity)
//   red_empty[i]   reducer -> consumers
template <typename T, int U, int NS, int C1, int C2, int C3, bool HAS_MULT, bool SR, int PD>
__global__ void __launch_bounds__(kChThreads, 1)
    timelag_chain_kernel(const T *__restrict__ mult, double *__restrict__ sums, double *__restrict__ cta_acc, long long F, long long A,
                         int D, int split, int upp, int4 pitch4, int blk_bytes, int nbuf, int w0, int ov, int nred, int halo, int align,
                         const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_first,
                         const __grid_constant__ CUtensorMap tmap_last, const __grid_constant__ CUtensorMap tmap_last4) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using V = typename ChVec<T>::type;
  constexpr int UE = 16 / (int)sizeof(T);
  constexpr int Q = 1 + NS + (C1 ? 1 : 0) + (C2 ? 1 : 0) + (C3 ? 1 : 0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  constexpr int RP = SR ? kChRedPitch : 1;
  unsigned char *ring = smem_raw;                                                          // [nbuf][blk_bytes]
  double *red = reinterpret_cast<double *>(ring + (size_t)nbuf * blk_bytes);               // [2][kChB][Q][RP]
  unsigned char *halo_buf = reinterpret_cast<unsigned char *>(red + 2 * kChB * Q * RP);    // [kChHalo][512 U] (halo only)
  uint64_t *full = reinterpret_cast<uint64_t *>(halo_buf + (halo ? kChHalo * 512 * U : 0)); // [kChNbufMax]
  uint64_t *empty = full + kChNbufMax;                                                     // [kChNbufMax]
  uint64_t *red_full = empty + kChNbufMax;                                                 // [2]
  uint64_t *red_empty = red_full + 2;                                                      // [2]
  uint64_t *halo_full = red_empty + 2;                                                     // [kChHalo]

  const long long my_atoms = (long long)blockIdx.x < A ? (A - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int n_my = (int)my_atoms * split;  // atom-major items: all pieces of an atom are consecutive items of this CTA
  const int nblk = (int)((F + kChB - 1) / kChB);

  if (tid == 0) {
    for (int s = 0; s < nbuf; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kChS);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&red_full[i], kChS);
      mbar_init(&red_empty[i], nred);
    }
    for (int i = 0; i < kChHalo; ++i) mbar_init(&halo_full[i], 1);
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == kChS) {
    // ===== producer: block g goes to buffer g % nbuf, free once block g - nbuf + 1 is consumed =======================
    int s = 0, ph = 0, g = 0;
    for (int k = 0; k < n_my; ++k) {
      const ChItem ci = ch_item<UE>(k, split, upp, D, align);
      const TlItem &it = ci.it;
      const int box_pitch = ci.kind == 0 ? pitch4.x : ci.kind == 1 ? pitch4.y : ci.kind == 2 ? pitch4.z : pitch4.w;
      const CUtensorMap *box_map = ci.kind == 0 ? &tmap : ci.kind == 1 ? &tmap_first : ci.kind == 2 ? &tmap_last : &tmap_last4;
      for (int b = 0; b < nblk; ++b, ++g) {
        if (g >= nbuf) mbar_wait(&empty[s], (uint32_t)(ph ^ 1));
        if (lane == 0) {
          // rows beyond F are zero-filled and still counted in the transaction bytes
          // (and so are the rows -ov .. -1 in front of an item's first block)
          mbar_expect_tx(&full[s], (uint32_t)(kChB + ov) * (uint32_t)box_pitch);
          tma_load_2d(ring + (size_t)s * blk_bytes, box_map, (int)it.w0, b * kChB - ov, &full[s]);
        }
        if (++s == nbuf) s = 0, ph ^= 1;
      }
    }
  } else if (warp > kChS + nred) {
    // (a reducer warp the launch does not use)
  } else if (warp > kChS) {
    // ===== reducers: add the lanes' partial sums, then the pieces of an atom ============================================
    const int rw = warp - (kChS + 1);  // each reducer warp takes its share of a block's (frame, sum) pairs
    int g = 0;
    const uint64_t pol = l2_policy_evict_last();
    for (int k = 0; k < n_my; ++k) {
      const TlItem it = tl_item<UE>(k, split, upp, D, true, false);
      double *srow = sums + (size_t)it.a * Q * (size_t)F;
      double *acc_row = cta_acc + (size_t)blockIdx.x * Q * (size_t)nblk * kChB;
      const bool first_piece = it.p == 0, last_piece = it.p == split - 1;
      for (int b = 0; b < nblk; ++b, ++g) {
        if constexpr (SR) {
          // lane <-> (frame, sum) pair of the block, 32 pairs per round: adds the 32 lanes' partial sums in a fixed
          // tree (16 pair sums, 4 chains of 4, 2 + 1), then the pieces of the atom
          constexpr int NP = kChB * Q, ROUNDS = (NP + 31) / 32;  // round rd belongs to reducer warp rd % nred
          double prev[ROUNDS];
          size_t where[ROUNDS];
          long long fls[ROUNDS];
          int qs[ROUNDS];
#pragma unroll
          for (int rd = 0; rd < ROUNDS; ++rd) {
            const int pi = (rd % nred == rw) ? rd * 32 + lane : NP, fb = pi / Q;
            qs[rd] = pi - fb * Q;
            fls[rd] = (long long)b * kChB + fb;
            where[rd] = (size_t)qs[rd] * nblk * kChB + (size_t)fls[rd];
            prev[rd] = (first_piece || pi >= NP) ? 0.0 : ld_acc(acc_row + where[rd], pol);
          }
          mbar_wait(&red_full[g & 1], (uint32_t)((g >> 1) & 1));
          const uint32_t rbase = smem_u32(red) + (uint32_t)(g & 1) * (uint32_t)(NP * kChRedPitch * 8);
#pragma unroll
          for (int rd = 0; rd < ROUNDS; ++rd) {
            const int pi = (rd % nred == rw) ? rd * 32 + lane : NP;
            if (pi < NP) {
              const uint32_t ra = rbase + (uint32_t)pi * (uint32_t)(kChRedPitch * 8);
              double e[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const double2 d = lds_d2(ra + 16u * i);
                e[i] = d.x + d.y;
              }
              double c[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) c[k] = ((e[k] + e[k + 4]) + e[k + 8]) + e[k + 12];
              double sum = (c[0] + c[1]) + (c[2] + c[3]);
              sum = first_piece ? sum : prev[rd] + sum;  // ((s_0 + s_1) + s_2) + ...: fixed order
              if (last_piece) {
                if (fls[rd] < F) srow[(size_t)qs[rd] * F + fls[rd]] = sum;
              } else {
                st_acc(acc_row + where[rd], sum, pol);
              }
            }
          }
        } else {
          const long long fl = (long long)b * kChB + lane;
          double prev[Q];
#pragma unroll
          for (int q = 0; q < Q; ++q)
            prev[q] = (first_piece || lane >= kChB || rw != 0) ? 0.0 : ld_acc(acc_row + (size_t)q * nblk * kChB + fl, pol);
          mbar_wait(&red_full[g & 1], (uint32_t)((g >> 1) & 1));
          if (lane < kChB && rw == 0) {  // (the butterfly leaves one value per frame and sum: one reducer warp does it)
            const double *r = red + ((size_t)(g & 1) * kChB + lane) * Q;
#pragma unroll
            for (int q = 0; q < Q; ++q) {
              const double sum = first_piece ? r[q] : prev[q] + r[q];  // ((s_0 + s_1) + s_2) + ...: fixed order
              if (last_piece) {
                if (fl < F) srow[(size_t)q * F + fl] = sum;
              } else {
                st_acc(acc_row + (size_t)q * nblk * kChB + fl, sum, pol);
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&red_empty[g & 1]);
      }
    }
  } else {
    // ===== consumers: warp <-> chain (frame mod 5), lane <-> units l + 32 c =============================================
    const uint32_t ring_addr = smem_u32(ring);
    const int red_q = ch_red_owner<Q>(lane);
    const bool red_writer = red_q >= 0 && (lane & 3) == 0;
    V h[kChH][U];
#pragma unroll
    for (int n = 0; n < kChH; ++n)
#pragma unroll
      for (int c = 0; c < U; ++c) {
        if constexpr (sizeof(T) == 8) h[n][c] = make_double2(0.0, 0.0);
        else h[n][c] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    const uint32_t col = 16u * (uint32_t)lane;  // column c of the lane: + 512 c
    int s = 0, ph = 0, g = 0;
    const uint32_t halo_addr = smem_u32(halo_buf);
    if (halo && warp == kChS - 1 && lane == 0) mbar_arrive(&halo_full[0]);  // (block 0 has no frame in front of it)

    // one block = kChSteps chain steps, history slots HALF * kChSteps + j (static)
    auto block = [&](auto half, const bool (&act)[U], const uint32_t (&cofs)[U], const V (&m)[U], const int pitch) {
      constexpr int HALF = decltype(half)::value;
      const uint32_t cur = ring_addr + (uint32_t)s * (uint32_t)blk_bytes;
      const uint32_t prv = ring_addr + (uint32_t)(s == 0 ? nbuf - 1 : s - 1) * (uint32_t)blk_bytes;
      mbar_wait(&full[s], (uint32_t)ph);
      if (g >= 2) mbar_wait(&red_empty[g & 1], (uint32_t)(((g >> 1) - 1) & 1));
      if (halo && warp == 0) mbar_wait(&halo_full[g % kChHalo], (uint32_t)((g / kChHalo) & 1));
      double *r = red + (size_t)(g & 1) * kChB * Q * RP;
      if constexpr (PD > 0) {
        // software pipeline over the kChSteps * U (step, column) tasks of the block: loads run PD columns ahead
        static_assert(SR, "the pipelined block hands its sums over through shared memory");
        constexpr int NT = kChSteps * U, NSLOT = PD + 1;
        using S = std::conditional_t<sizeof(T) == 8, double, float>;
        V xq[NSLOT], pq[NSLOT];
        S sacc[Q][2];
        auto load = [&](auto tt) {
          constexpr int TT = decltype(tt)::value, J = TT / U, C = TT % U;
          const int fl = kChS * J + warp + ov;  // row of the frame in the block's buffer
          const int pl = fl - w0;               // its ring-window partner: this buffer or the previous one (w0 <= 25)
          ch_lds(xq[TT % NSLOT], cur + (uint32_t)fl * (uint32_t)pitch + col + 512u * C, act[C]);
          if (NS)
            ch_lds(pq[TT % NSLOT], (pl >= 0 ? cur + (uint32_t)pl * (uint32_t)pitch : prv + (uint32_t)(pl + kChB) * (uint32_t)pitch) + col + 512u * C, act[C]);
        };
        ch_static_for(load, std::make_integer_sequence<int, (PD < NT ? PD : NT)>{});
        auto task = [&](auto tt) {
          constexpr int TT = decltype(tt)::value, J = TT / U, C = TT % U;
          if constexpr (TT + PD < NT) load(std::integral_constant<int, TT + PD>{});
          if constexpr (C == 0) {
#pragma unroll
            for (int q = 0; q < Q; ++q) sacc[q][0] = sacc[q][1] = (S)0;
          }
          ch_col<HALF * kChSteps + J, C, U, NS, C1, C2, C3, Q>(xq[TT % NSLOT], pq[TT % NSLOT], m[C], h, sacc);
          if constexpr (C == U - 1) {
            const int fl = kChS * J + warp;
#pragma unroll
            for (int q = 0; q < Q; ++q) r[(fl * Q + q) * kChRedPitch + lane] = (double)sacc[q][0] + (double)sacc[q][1];
          }
        };
        ch_static_for(task, std::make_integer_sequence<int, NT>{});
      } else {
      auto step = [&](auto jj) {
        constexpr int J = decltype(jj)::value;
        const int fl = kChS * J + warp;  // frame within the block
        const uint32_t x_addr = cur + (uint32_t)(fl + ov) * (uint32_t)pitch;
        // ring-window partner: row fl + ov - w0 of this buffer, or of the previous one (w0 <= 25, ov = 0; the first
        // frames of an item read the previous item's last frames there: results for f < w0 are never used)
        const int pl = fl + ov - w0;
        uint32_t p_addr = pl >= 0 ? cur + (uint32_t)pl * (uint32_t)pitch : prv + (uint32_t)(pl + kChB) * (uint32_t)pitch;
        uint32_t halo_st = 0;
        if constexpr (J == 0) {  // (halo: w0 == 1) the frame in front of the block comes from chain 4's registers
          if (halo && warp == 0) p_addr = halo_addr + (uint32_t)(g % kChHalo) * (512u * U);
        }
        if constexpr (J == kChSteps - 1) {
          if (halo && warp == kChS - 1) halo_st = halo_addr + (uint32_t)((g + 1) % kChHalo) * (512u * U);
        }
        double v[Q];
        if constexpr (sizeof(T) == 8)
          ch_step_f64<HALF * kChSteps + J, U, NS, C1, C2, C3, Q>(x_addr, p_addr, col, cofs, m, h, v, halo_st);
        else
          ch_step_f32<HALF * kChSteps + J, U, NS, C1, C2, C3, Q>(x_addr, p_addr, col, act, m, h, v, halo_st);
        if constexpr (SR) {
#pragma unroll
          for (int q = 0; q < Q; ++q) r[(fl * Q + q) * kChRedPitch + lane] = v[q];
        } else {
          const double tot = ch_reduce<Q>(v, lane);
          if (red_writer) r[fl * Q + red_q] = tot;
        }
      };
      static_assert(kChSteps == 5, "five chain steps per block are spelled out");
      step(std::integral_constant<int, 0>{});
      step(std::integral_constant<int, 1>{});
      step(std::integral_constant<int, 2>{});
      step(std::integral_constant<int, 3>{});
      step(std::integral_constant<int, 4>{});
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&red_full[g & 1]);
        if (halo && warp == kChS - 1) mbar_arrive(&halo_full[(g + 1) % kChHalo]);
        if (ov || NS == 0 || halo)
          mbar_arrive(&empty[s]);  // (ov == w0, the halo, or no ring window at all: nothing of this buffer is needed again)
        else if (g >= 1)
          mbar_arrive(&empty[s == 0 ? nbuf - 1 : s - 1]);
      }
      if (++s == nbuf) s = 0, ph ^= 1;
      ++g;
    };

    for (int k = 0; k < n_my; ++k) {
      const ChItem ci = ch_item<UE>(k, split, upp, D, align);
      const TlItem &it = ci.it;
      bool act[U];
      V m[U];
#pragma unroll
      for (int c = 0; c < U; ++c) {
        act[c] = lane + 32 * c < it.punits;
        if constexpr (sizeof(T) == 8) {
          m[c] = make_double2(act[c] ? 1.0 : 0.0, act[c] ? 1.0 : 0.0);
          if (HAS_MULT && act[c]) m[c] = *reinterpret_cast<const double2 *>(mult + it.e0 + (lane + 32 * c) * UE);
        } else {
          const float one = act[c] ? 1.f : 0.f;
          m[c] = make_float4(one, one, one, one);
          if (HAS_MULT && act[c]) m[c] = *reinterpret_cast<const float4 *>(mult + it.e0 + (lane + 32 * c) * UE);
        }
      }
      // (the first block of an item reads its ring-window partners from the previous ITEM's last block with this
      // item's pitch: garbage either way, and never used)
      const int pitch = ci.kind == 0 ? pitch4.x : ci.kind == 1 ? pitch4.y : ci.kind == 2 ? pitch4.z : pitch4.w;
      uint32_t cofs[U];  // byte offset of the lane's unit in column c (clamped to the last unit of the piece)
#pragma unroll
      for (int c = 0; c < U; ++c) cofs[c] = 16u * (uint32_t)min(lane + 32 * c, it.punits - 1);
      for (int b = 0; b < nblk; b += 2) {
        block(std::integral_constant<int, 0>{}, act, cofs, m, pitch);
        if (b + 1 < nblk) block(std::integral_constant<int, 1>{}, act, cofs, m, pitch);
      }
    }
  }
}

}  // namespace soapb
