// timelag.cu -- K3/K4: per-atom time-lagged SOAP distances (timeSOAP), all requested windows in
// ONE pass over the trajectory, optionally straight from the COMPRESSED power spectrum.
//
// Reference semantics: timeSOAP (src/SOAPify/analysis.py:13-69)
//     t_w[f-w][a] = dist(X[f][a][:], X[f-w][a][:])
// with dist one of src/SOAPify/distances.py; deltaTimedSOAP = numpy.diff(t.T, axis=-1)
// (analysis.py:67).  With `mult` = multiplicity of every compressed slot in the gather table
// of fillSOAPVectorFromdscribe (utils.py:212-268) the weighted dots  sum_i m_i x_i y_i  ARE the
// dots of the expanded vectors, so expansion and normalisation never touch HBM.
//
// Layout in HBM: X[F][A][D], row (f,a) contiguous.  The lag partner of a row is A*D elements
// away, so reuse must come from on-chip memory: one CTA owns (atom a, piece p of the row) and
// WALKS THE FRAMES, keeping the last R frames of its piece in a shared-memory ring filled by the
// TMA engine (one 1-D cp.async.bulk per frame, 32 frames per mbarrier).  Splitting the row
// into pieces is what lets a 50-frame history of a 9.8 kB row fit in 227 kB.
//
// Thread mapping: LANE = FRAME.  A warp takes 32 consecutive frames, every lane walks the
// elements of its own frame (128-bit LDS, ring slots are 16*odd bytes apart -> conflict free)
// and multiplies with the lagged frames' elements held (in the same ring) by "earlier" slots.
// No cross-lane reduction is ever needed: each lane ends with the complete partial sums of its
// frame; the 8 warps split the elements and are summed once per 32-frame block.
// Partial sums (uu, uv_w) are added up over the pieces of an atom in a per-CTA accumulator that lives in L2,
// the totals go to sums[a][q][f]; a small finalize kernel (which used to add the pieces itself, in
// fixed order (deterministic), applies the distance epilogue and writes t_w[f-w][a].
#include "tc05.cuh"

#include <stdlib.h>
#include <string.h>
#include <type_traits>
#include <utility>

namespace soapb {

constexpr int kBlk = 32;     // frames per block == lanes per warp
constexpr int kTlWarps = 8;  // consumer warps (two per SM sub-partition)
constexpr int kChainAlignDefault = 1;  // SOAP_TL_CHAIN_ALIGN default: on (+1.4 % where the kernel is HBM bound, level where it is SM bound)
constexpr int kChainHaloDefault = 1;  // SOAP_TL_CHAIN_HALO default: on (3-5 % faster than overlapping boxes, interleaved A/B)
constexpr int kChainDefault = -1;  // SOAP_TL_CHAIN default (timelag_chain.cuh): -1 = float64 data only (measured)
constexpr int kPairDefault = -1;  // SOAP_TL_PAIR default (timelag_pair.cuh): -1 = float32 data only (measured)

struct TimelagPlan {
  int unit_elems, n_units, split, upp, stride, R, ahead, Q, max_lag, cw, mbuf;
  size_t smem;
};

static size_t tl_smem_bytes(int R, int stride, int upp, int Q, int cw, int mbuf) {
  return (size_t)R * stride + (size_t)mbuf * upp * 16 + (size_t)cw * Q * 32 * 8 + (size_t)(2 * (R / kBlk) + 2) * 8 + 64;
}

// upper bound on the persistent grid (CTAs resident at once); sizes the per-CTA accumulators in scratch
static long long tl_max_grid() { return (long long)num_sms() * 4; }

// a CTA owns whole atoms (and sums their pieces on chip) when there are enough atoms to keep every SM busy
// with little imbalance; with fewer atoms the (atom, piece) items are dealt round robin
static bool tl_atom_major(int64_t n_atoms) { return n_atoms >= 4ll * num_sms(); }

static const char *g_tl_last_kernel = "";

static int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

static bool make_plan(int64_t F, int D, int elem_size, const int *windows, int nw, TimelagPlan &pl) {
  pl.unit_elems = 16 / elem_size;
  pl.n_units = (D + pl.unit_elems - 1) / pl.unit_elems;
  pl.Q = 1 + nw;
  pl.max_lag = 0;
  for (int k = 0; k < nw; ++k) pl.max_lag = windows[k] > pl.max_lag ? windows[k] : pl.max_lag;
  const size_t budget = (size_t)max_smem_optin() - 1024;
  const int force_split = env_int("SOAP_TL_SPLIT", 0);    // tuning knobs (benchmarks only)
  const int force_ahead = env_int("SOAP_TL_AHEAD", 0);
  pl.cw = env_int("SOAP_TL_WARPS", kTlWarps);
  for (int split = force_split > 0 ? force_split : 1; split <= pl.n_units; ++split) {
    // units per piece: odd (ring slots 16*odd bytes apart -> conflict-free 128-bit LDS) and <= 127
    // (one tensor-map box row holds at most 256 8-byte elements)
    const int upp = ((pl.n_units + split - 1) / split) | 1;
    if (upp > 127) continue;
    const int real_split = (pl.n_units + upp - 1) / upp;  // no empty pieces
    const int stride = 16 * upp;
    for (int ahead = force_ahead > 0 ? force_ahead : 3; ahead >= 1; --ahead) {
      // ring = blocks still needed as lag partners + the block being consumed + `ahead` in flight
      const int R = kBlk * ((pl.max_lag + kBlk - 1) / kBlk + 1 + ahead);
      // multiplicity buffers: item k's may only be overwritten once item k - mbuf is consumed, which the
      // producer knows (from the ring's own hand-shake) when mbuf - 1 items are at least `ahead` blocks
      const int64_t nblk = (F + kBlk - 1) / kBlk;
      const int mbuf = nblk >= ahead ? 2 : 4;
      const size_t smem = tl_smem_bytes(R, stride, upp, pl.Q, pl.cw, mbuf);
      if (smem <= budget) {
        pl.mbuf = mbuf;
        pl.split = real_split;
        pl.upp = upp;
        pl.stride = stride;
        pl.R = R;
        pl.ahead = ahead;
        pl.smem = smem;
        return true;
      }
      if (force_ahead > 0) break;
    }
  }
  return false;
}

// one element global -> shared; `valid == false` writes zeros instead (src-size 0)
template <size_t BYTES>
__device__ __forceinline__ void cp_async_elem(uint32_t dst, const void *src, bool valid) {
  const uint32_t n = valid ? (uint32_t)BYTES : 0u;
  if constexpr (BYTES == 8)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t *bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// U units (16 bytes each) of the lane's frame against the NW lagged frames.  All loads are issued
// before the first use so that one warp keeps U*(NW+2) shared-memory requests in flight.
// fp64: exact products accumulated by DFMA, separate chains for the two halves of a unit.
template <int NW, int U, int EUCLID, bool HM, bool CLEAN = false>
__device__ __forceinline__ void accum_units_f64(uint32_t m_addr, uint32_t x_addr, const uint32_t (&y_addr)[NW],
                                                uint32_t off, uint32_t ustep, double (&acc)[2 * (NW + 1)],
                                                const double *sc = nullptr) {
  double2 m[U], x[U], y[NW][U];
#pragma unroll
  for (int j = 0; j < U; ++j) {
    m[j] = HM ? lds_d2(m_addr + off + j * ustep) : make_double2(1.0, 1.0);
    x[j] = lds_d2(x_addr + off + j * ustep);
#pragma unroll
    for (int k = 0; k < NW; ++k) y[k][j] = lds_d2(y_addr[k] + off + j * ustep);
  }
  if (CLEAN) {  // foreign elements (multiplicity 0) of a row's first / last unit must not leak as 0 * NaN
#pragma unroll
    for (int j = 0; j < U; ++j) {
      x[j].x = m[j].x != 0.0 ? x[j].x : 0.0;
      x[j].y = m[j].y != 0.0 ? x[j].y : 0.0;
#pragma unroll
      for (int k = 0; k < NW; ++k) {
        y[k][j].x = m[j].x != 0.0 ? y[k][j].x : 0.0;
        y[k][j].y = m[j].y != 0.0 ? y[k][j].y : 0.0;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < U; ++j) {
    const double xm0 = x[j].x * m[j].x, xm1 = x[j].y * m[j].y;
    acc[0] = fma(xm0, x[j].x, acc[0]);
    acc[1] = fma(xm1, x[j].y, acc[1]);
#pragma unroll
    for (int k = 0; k < NW; ++k) {
      if (EUCLID) {
        // EUCLID == 2: the rows are normalised on the fly, sc[0] = 1/|x|, sc[1 + k] = 1/|y_k| (zero norm -> 1)
        const double d0 = EUCLID == 2 ? fma(x[j].x, sc[0], -(y[k][j].x * sc[1 + k])) : x[j].x - y[k][j].x;
        const double d1 = EUCLID == 2 ? fma(x[j].y, sc[0], -(y[k][j].y * sc[1 + k])) : x[j].y - y[k][j].y;
        acc[2 * k + 2] = fma(m[j].x * d0, d0, acc[2 * k + 2]);
        acc[2 * k + 3] = fma(m[j].y * d1, d1, acc[2 * k + 3]);
      } else {
        acc[2 * k + 2] = fma(xm0, y[k][j].x, acc[2 * k + 2]);
        acc[2 * k + 3] = fma(xm1, y[k][j].y, acc[2 * k + 3]);
      }
    }
  }
}

// fp32 data: products and a chain of 4*U terms in fp32, chain sums promoted to double (one
// conversion per 4*U products; the fp32 rounding is bounded to one short chain).
template <int NW, int U, int EUCLID, bool HM, bool CLEAN = false>
__device__ __forceinline__ void accum_units_f32(uint32_t m_addr, uint32_t x_addr, const uint32_t (&y_addr)[NW],
                                                uint32_t off, uint32_t ustep, double (&acc)[2 * (NW + 1)],
                                                const double *sc = nullptr) {
  float4 m[U], x[U], y[NW][U];
#pragma unroll
  for (int j = 0; j < U; ++j) {
    m[j] = HM ? lds_f4(m_addr + off + j * ustep) : make_float4(1.f, 1.f, 1.f, 1.f);
    x[j] = lds_f4(x_addr + off + j * ustep);
#pragma unroll
    for (int k = 0; k < NW; ++k) y[k][j] = lds_f4(y_addr[k] + off + j * ustep);
  }
  if (CLEAN) {  // foreign elements (multiplicity 0) of a row's first / last unit must not leak as 0 * NaN
#pragma unroll
    for (int j = 0; j < U; ++j) {
      x[j].x = m[j].x != 0.f ? x[j].x : 0.f;
      x[j].y = m[j].y != 0.f ? x[j].y : 0.f;
      x[j].z = m[j].z != 0.f ? x[j].z : 0.f;
      x[j].w = m[j].w != 0.f ? x[j].w : 0.f;
#pragma unroll
      for (int k = 0; k < NW; ++k) {
        y[k][j].x = m[j].x != 0.f ? y[k][j].x : 0.f;
        y[k][j].y = m[j].y != 0.f ? y[k][j].y : 0.f;
        y[k][j].z = m[j].z != 0.f ? y[k][j].z : 0.f;
        y[k][j].w = m[j].w != 0.f ? y[k][j].w : 0.f;
      }
    }
  }
  float c0 = 0.f, c1 = 0.f;
#pragma unroll
  for (int j = 0; j < U; ++j) {
    c0 = fmaf(x[j].x * m[j].x, x[j].x, c0);
    c1 = fmaf(x[j].y * m[j].y, x[j].y, c1);
    c0 = fmaf(x[j].z * m[j].z, x[j].z, c0);
    c1 = fmaf(x[j].w * m[j].w, x[j].w, c1);
  }
  acc[0] += (double)c0;
  acc[1] += (double)c1;
#pragma unroll
  for (int k = 0; k < NW; ++k) {
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int j = 0; j < U; ++j) {
      if (EUCLID) {
        float d0, d1, d2, d3;
        if (EUCLID == 2) {  // normalised on the fly (float32 arithmetic, like the reference on float32 data)
          const float sx = (float)sc[0], sy = (float)sc[1 + k];
          d0 = fmaf(x[j].x, sx, -(y[k][j].x * sy));
          d1 = fmaf(x[j].y, sx, -(y[k][j].y * sy));
          d2 = fmaf(x[j].z, sx, -(y[k][j].z * sy));
          d3 = fmaf(x[j].w, sx, -(y[k][j].w * sy));
        } else {
          d0 = x[j].x - y[k][j].x, d1 = x[j].y - y[k][j].y, d2 = x[j].z - y[k][j].z, d3 = x[j].w - y[k][j].w;
        }
        s0 = fmaf(m[j].x * d0, d0, s0);
        s1 = fmaf(m[j].y * d1, d1, s1);
        s0 = fmaf(m[j].z * d2, d2, s0);
        s1 = fmaf(m[j].w * d3, d3, s1);
      } else {
        s0 = fmaf(x[j].x * m[j].x, y[k][j].x, s0);
        s1 = fmaf(x[j].y * m[j].y, y[k][j].y, s1);
        s0 = fmaf(x[j].z * m[j].z, y[k][j].z, s0);
        s1 = fmaf(x[j].w * m[j].w, y[k][j].w, s1);
      }
    }
    acc[2 * k + 2] += (double)s0;
    acc[2 * k + 3] += (double)s1;
  }
}

template <typename T, int NW, int U, int EUCLID, bool HM, bool CLEAN = false>
__device__ __forceinline__ void accum_units(uint32_t m_addr, uint32_t x_addr, const uint32_t (&y_addr)[NW],
                                            uint32_t off, uint32_t ustep, double (&acc)[2 * (NW + 1)],
                                            const double *sc = nullptr) {
  if constexpr (sizeof(T) == 8)
    accum_units_f64<NW, U, EUCLID, HM, CLEAN>(m_addr, x_addr, y_addr, off, ustep, acc, sc);
  else
    accum_units_f32<NW, U, EUCLID, HM, CLEAN>(m_addr, x_addr, y_addr, off, ustep, acc, sc);
}

// Warp roles: warps 0..CW-1 consume (lane = frame), warp CW produces (TMA / cp.async), warp CW+1 reduces
// the consumers' partial sums of a block and writes them to scratch.  All hand-offs are
// mbarriers; there is no __syncthreads after the prologue, so no warp ever convoys behind another.
//   full[s]      producer -> consumers   block in ring slot-group s has landed
//   empty[s]     consumers -> producer   block s consumed as "current" frames (one arrival per warp)
//   red_full     consumers -> reducer    partial sums of a block are in red (one arrival per warp)
//   red_empty    reducer -> consumers    red may be overwritten
// The kernel is PERSISTENT: a CTA walks the work items i = blockIdx.x, + gridDim.x, ... (item = atom a,
// piece p) as one continuous stream of 32-frame blocks, so the producer is already fetching the next
// item's first frames while the consumers finish the current one (no per-item pipeline fill).
struct TlItem {
  long long a;
  int p, e0, pe, punits;
  int lead;       // shifted rows: elements of the PREVIOUS row in front of this one (the box starts 16-byte aligned)
  long long w0;   // first 8-byte word of the piece in the frame row (TMA coordinate)
};

// per-CTA accumulator traffic: keep the lines in L2 (evict_last) and out of L1 -- between a store and the
// load of the next piece ~300 MB of trajectory stream through L2 and would otherwise push them out to DRAM
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void st_acc(double *p, double v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ double ld_acc(const double *p, uint64_t pol) {
  double v;
  asm volatile("ld.global.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol) : "memory");
  return v;
}

// item j of a CTA.  atom_major: its atoms are blockIdx.x, + gridDim.x, ... and ALL pieces of an atom are
// consecutive items of the same CTA, so that the reducer can add the pieces up before anything goes to DRAM.
// Otherwise (few atoms: fewer atoms than would keep every SM busy) items (atom, piece) are dealt round robin
// and every piece writes its own partial sums, which the finalize kernel adds.
// `shifted` (rows that are not 16-byte multiples): the TMA engine wants 16-byte aligned box rows (a box that starts
// 4 or 8 bytes into a unit raises an illegal-instruction fault, scripts/micro/tma_align.cu), so the window of an atom
// whose row starts inside a unit begins at the unit boundary below it; its first unit then carries `lead` foreign
// elements, and / or its last unit ends with foreign elements.  The multiplicity table has one variant per lead
// (0 .. UE-1 elements), zero on the foreign elements, and the consumers clean those units (CLEAN).
template <int UE>
__device__ __forceinline__ TlItem tl_item(int j, int split, int upp, int D, bool atom_major, bool shifted = false) {
  TlItem it;
  if (atom_major) {
    const int k = j / split;
    it.p = j - k * split;
    it.a = (long long)blockIdx.x + (long long)k * gridDim.x;
  } else {
    const long long i = (long long)blockIdx.x + (long long)j * gridDim.x;
    it.a = i / split;
    it.p = (int)(i - it.a * split);
  }
  constexpr int ES = 16 / UE;
  const long long s0 = it.a * D;                 // first element of the row
  it.lead = shifted ? (int)(s0 % UE) : 0;        // elements between the 16-byte boundary below the row and the row
  it.e0 = it.p * upp * UE;                       // relative to the (aligned) window
  it.pe = min(it.lead + D - it.e0, upp * UE);    // elements of this piece; <= 0 for an empty last piece
  it.punits = it.pe > 0 ? (it.pe + UE - 1) / UE : 0;
  it.w0 = (s0 - it.lead + it.e0) * ES / 8;       // (a whole number of units: an even word)
  return it;
}

template <typename T, int NW, int EUCLID, bool HAS_MULT, int CW, int UB>
__global__ void __launch_bounds__((CW + 2) * 32, 1)
    timelag_kernel(const T *__restrict__ X, const T *__restrict__ mult, double *__restrict__ sums,
                   double *__restrict__ cta_acc, const double *__restrict__ inv_norm,
                   long long F, long long A, int D, int split, int upp, int stride, int R, int lag_blocks,
                   int mbuf, int4 win4, int bulk_ok, int mult_pitch, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int UE = 16 / (int)sizeof(T);
  constexpr int Q = NW + 1;
  constexpr int kTlThreads = (CW + 2) * 32;  // consumers + producer warp + reducer warp
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int win[4] = {win4.x, win4.y, win4.z, win4.w};
  const int nslots = R / kBlk;

  unsigned char *ring = smem_raw;
  unsigned char *msm = ring + (size_t)R * stride;                          // [mbuf][upp] units
  double *red = reinterpret_cast<double *>(msm + (size_t)mbuf * upp * 16);  // [CW][Q][32]
  uint64_t *full = reinterpret_cast<uint64_t *>(red + CW * Q * 32);  // [nslots]
  uint64_t *empty = full + nslots;                                        // [nslots]
  uint64_t *red_full = empty + nslots;                                    // [1]
  uint64_t *red_empty = red_full + 1;                                     // [1]

  const bool atom_major = cta_acc != nullptr;
  const long long my_atoms = (long long)blockIdx.x < A ? (A - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const long long n_items = A * split;
  const int n_my = atom_major ? (int)my_atoms * split
                              : ((long long)blockIdx.x < n_items ? (int)((n_items - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0);
  const int nblk = (int)((F + kBlk - 1) / kBlk);

  // prologue (all warps): the ring is zeroed when rows are not 16-byte multiples (the tail unit of a
  // row is then partial and its padding must read as zero)
  if (!(bulk_ok & 1)) {
    const int n16 = (int)(((size_t)R * stride + (size_t)mbuf * upp * 16) / 16);
    for (int i = tid; i < n16; i += kTlThreads) reinterpret_cast<int4 *>(ring)[i] = make_int4(0, 0, 0, 0);
  }
  if (tid == 0) {
    for (int s = 0; s < nslots; ++s) {
      mbar_init(&full[s], (bulk_ok & 1) ? 1 : 32);
      mbar_init(&empty[s], CW);
    }
    mbar_init(red_full, CW);
    mbar_init(red_empty, 1);
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == CW) {
    // ===== producer ==========================================================================
    const size_t fstride = (size_t)A * D;  // elements between consecutive frames
    const uint32_t ring_addr = smem_u32(ring), msm_addr = smem_u32(msm);
    // loading block g overwrites the frames of block g - nslots, last needed (as lag partners) by
    // block j = g - back: wait until the consumers are done with block j
    const int back = nslots - lag_blocks;
    int s = 0, js = 0, jph = 0, g = 0;
    for (int k = 0; k < n_my; ++k) {
      const TlItem it = tl_item<UE>(k, split, upp, D, atom_major, (bulk_ok & 32) != 0);
      const T *xa = X + (size_t)it.a * D + it.e0;
      const int mslot = k & (mbuf - 1);
      // shifted rows: one multiplicity table per lead ([UE][mult_pitch] elements)
      const T *mrow = mult + ((bulk_ok & 32) ? it.lead * mult_pitch : 0) + it.e0;
      for (int b = 0; b < nblk; ++b, ++g) {
        if (g >= back) {
          mbar_wait(&empty[js], (uint32_t)jph);
          if (++js == nslots) js = 0, jph ^= 1;
        }
        const long long f0 = (long long)b * kBlk;
        if (bulk_ok & 2) {  // diagnostic: no copies
          if (lane == 0) mbar_arrive(&full[s]);
        } else if (bulk_ok & 1) {
          // ONE tensor-map box = 32 frames x this piece (rows beyond F and columns beyond the tensor are
          // zero-filled and still counted in the transaction bytes); the item's multiplicities ride on
          // the barrier of its first block
          if (lane == 0) {
            const uint32_t mbytes = (HAS_MULT && b == 0) ? (uint32_t)it.punits * 16u : 0u;  // (padded with zeros by the host side)
            mbar_expect_tx(&full[s], (uint32_t)kBlk * (uint32_t)stride + mbytes);
            tma_load_2d(ring + (size_t)s * kBlk * stride, &tmap, (int)it.w0, (int)f0, &full[s]);
            if (mbytes) bulk_g2s(msm + (size_t)mslot * upp * 16, mrow, mbytes, &full[s]);
          }
        } else {
          const int nf = (int)min((long long)kBlk, F - f0);
          const int padded = it.punits * UE;  // the tail of a partial last unit is written as zeros
          for (int fr = 0; fr < nf; ++fr) {
            const T *src = xa + (size_t)(f0 + fr) * fstride;
            const uint32_t dst = ring_addr + (uint32_t)(s * kBlk + fr) * stride;
            for (int e = lane; e < padded; e += 32)
              cp_async_elem<sizeof(T)>(dst + e * (uint32_t)sizeof(T), src + min(e, it.pe - 1), e < it.pe);
          }
          if (HAS_MULT && b == 0) {
            const uint32_t dst = msm_addr + (uint32_t)mslot * upp * 16;
            for (int e = lane; e < padded; e += 32)
              cp_async_elem<sizeof(T)>(dst + e * (uint32_t)sizeof(T), mult + it.e0 + min(e, it.pe - 1), e < it.pe);
          }
          cp_async_arrive_noinc(&full[s]);  // every lane: arrives once its own copies have landed
        }
        if (++s == nslots) s = 0;
      }
    }
  } else if (warp == CW + 1) {
    // ===== reducer ===========================================================================
    int g = 0;
    const uint64_t pol = l2_policy_evict_last();
    for (int k = 0; k < n_my; ++k) {
      const TlItem it = tl_item<UE>(k, split, upp, D, atom_major, (bulk_ok & 32) != 0);
      // pieces 0 .. split-2 leave their running sums in this CTA's accumulator (read back by the same lane for
      // the next piece: L2 traffic only), the last piece writes the totals of the atom
      double *srow = atom_major ? sums + (size_t)it.a * Q * (size_t)F : sums + ((size_t)it.a * split + it.p) * Q * (size_t)F;
      double *acc_row = cta_acc + (size_t)blockIdx.x * Q * (size_t)nblk * kBlk;
      const bool first_piece = !atom_major || it.p == 0, last_piece = !atom_major || it.p == split - 1;
      for (int b = 0; b < nblk; ++b, ++g) {
        const long long fl = (long long)b * kBlk + lane;
        double prev[Q];
#pragma unroll
        for (int q = 0; q < Q; ++q) prev[q] = first_piece ? 0.0 : ld_acc(acc_row + (size_t)q * nblk * kBlk + fl, pol);
        mbar_wait(red_full, (uint32_t)(g & 1));
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          double sum = 0.0;
#pragma unroll
          for (int w = 0; w < CW; ++w) sum += red[(w * Q + q) * 32 + lane];
          sum = first_piece ? sum : prev[q] + sum;  // ((s_0 + s_1) + s_2) + ...: fixed order
          if (last_piece) {
            if (fl < F) srow[(size_t)q * F + fl] = sum;
          } else {
            st_acc(acc_row + (size_t)q * nblk * kBlk + fl, sum, pol);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(red_empty);
      }
    }
  } else {
    // ===== consumers: lane <-> frame ============================================================
    const uint32_t ring_addr = smem_u32(ring), msm_addr = smem_u32(msm);
    const uint32_t ustep = 16u * CW;  // this warp's units are u = warp, warp + CW, ...
    const uint32_t off0 = 16u * warp;
    int wmod[NW];
#pragma unroll
    for (int k = 0; k < NW; ++k) wmod[k] = win[k] % R;
    int s = 0, ph = 0, g = 0;
    for (int k = 0; k < n_my; ++k) {
      const TlItem it = tl_item<UE>(k, split, upp, D, atom_major, (bulk_ok & 32) != 0);
      int my_units = it.punits > warp ? (it.punits - warp + CW - 1) / CW : 0;
      // shifted rows: the unit that starts with the previous row's elements (first unit of piece 0 = warp 0's
      // first) and the unit that ends with the next row's (the last one) are taken separately, cleaned
      const bool head = HAS_MULT && (bulk_ok & 32) && it.lead != 0 && it.p == 0 && warp == 0 && my_units > 0;
      const bool tail = HAS_MULT && (bulk_ok & 32) && it.pe > 0 && (it.pe % UE) != 0 && (it.punits - 1) % CW == warp &&
                        !(head && my_units == 1);
      if (tail) --my_units;
      const uint32_t m_addr = msm_addr + (uint32_t)(k & (mbuf - 1)) * upp * 16;
      for (int b = 0; b < nblk; ++b, ++g) {
        const int slot = s * kBlk + lane;  // ring slot of this lane's frame (R is a multiple of 32)
        const uint32_t x_addr = ring_addr + (uint32_t)slot * stride;
        uint32_t y_addr[NW];
#pragma unroll
        for (int q = 0; q < NW; ++q) {
          int ys = slot - wmod[q];
          if (ys < 0) ys += R;
          y_addr[q] = ring_addr + (uint32_t)ys * stride;  // frames f < w read stale data: result unused
        }
        double acc[2 * Q];
#pragma unroll
        for (int q = 0; q < 2 * Q; ++q) acc[q] = 0.0;
        double sc[Q];  // EUCLID == 2: reciprocal norms of this lane's frame and of its lag partners
        if constexpr (EUCLID == 2) {
          const long long f = (long long)b * kBlk + lane;
          const double *ia = inv_norm + (size_t)it.a * (size_t)F;
          sc[0] = f < F ? ia[f] : 0.0;
#pragma unroll
          for (int q = 0; q < NW; ++q) sc[1 + q] = (f < F && f >= win[q]) ? ia[f - win[q]] : 0.0;
        }
        mbar_wait(&full[s], (uint32_t)ph);
        int n = 0;
        // (masking the lanes past the last frame -- 24 of 32 in the last block at F = 1000 -- measured 3 % SLOWER:
        // the divergent loop costs more than the wavefronts it saves)
        if (head) {
          accum_units<T, NW, 1, EUCLID, HAS_MULT, true>(m_addr, x_addr, y_addr, off0, ustep, acc, sc);
          n = 1;
        }
        if (bulk_ok & 4) n = my_units;  // diagnostic: no accumulation
        for (; n + UB <= my_units; n += UB)
          accum_units<T, NW, UB, EUCLID, HAS_MULT>(m_addr, x_addr, y_addr, off0 + n * ustep, ustep, acc, sc);
        if constexpr (UB > 4) {
          if (n + 4 <= my_units) {
            accum_units<T, NW, 4, EUCLID, HAS_MULT>(m_addr, x_addr, y_addr, off0 + n * ustep, ustep, acc, sc);
            n += 4;
          }
        }
        if constexpr (UB > 2) {
          if (n + 2 <= my_units) {
            accum_units<T, NW, 2, EUCLID, HAS_MULT>(m_addr, x_addr, y_addr, off0 + n * ustep, ustep, acc, sc);
            n += 2;
          }
        }
        if (n < my_units)
          accum_units<T, NW, 1, EUCLID, HAS_MULT>(m_addr, x_addr, y_addr, off0 + n * ustep, ustep, acc, sc);
        if (tail && !(bulk_ok & 4))
          accum_units<T, NW, 1, EUCLID, HAS_MULT, true>(m_addr, x_addr, y_addr, off0 + my_units * ustep, ustep, acc, sc);
        // hand the partial sums to the reducer (it is ~10x faster than a block: this wait is free)
        if (g >= 1) mbar_wait(red_empty, (uint32_t)((g - 1) & 1));
        double *r = red + (size_t)warp * Q * 32;
#pragma unroll
        for (int q = 0; q < Q; ++q) r[q * 32 + lane] = acc[2 * q] + acc[2 * q + 1];
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(red_full);
          mbar_arrive(&empty[s]);
        }
        if (++s == nslots) s = 0, ph ^= 1;
      }
    }
  }
}

// sums the pieces, applies the distance epilogue, writes t_w[f-w][a] through a 32x32 transpose
template <int NW>
__global__ void __launch_bounds__(256)
    timelag_finalize_kernel(const double *__restrict__ scratch, double *__restrict__ t, long long o0,
                            long long o1, long long o2, long long o3, long long F, long long A,
                            int split, int4 win4, int kind, int power) {
  constexpr int Q = NW + 1;
  __shared__ double tile[NW][32][33];
  const int win[4] = {win4.x, win4.y, win4.z, win4.w};
  const long long off[4] = {o0, o1, o2, o3};
  const long long f0 = (long long)blockIdx.x * 32, a0 = (long long)blockIdx.y * 32;
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  for (int i = 0; i < 4; ++i) {
    const int al = ly + 8 * i;
    const long long a = a0 + al, f = f0 + lx;
    if (a < A && f < F) {
      const double *base = scratch + (size_t)a * split * Q * (size_t)F;
      // all loads of a thread are independent: issued before the (fixed-order) sums so that a thread keeps
      // Q + NW of them in flight per piece instead of one
      double uu = 0.0, uv[NW], vv[NW];
#pragma unroll
      for (int k = 0; k < NW; ++k) uv[k] = vv[k] = 0.0;
#pragma unroll 2
      for (int p = 0; p < split; ++p) {
        const double *bp = base + ((size_t)p * Q) * F;
        const double u = bp[f];
        double a[NW], b[NW];
#pragma unroll
        for (int k = 0; k < NW; ++k) {
          const bool on = f >= win[k];
          a[k] = on ? bp[(size_t)(1 + k) * F + f] : 0.0;
          b[k] = on ? bp[f - win[k]] : 0.0;
        }
        uu += u;
#pragma unroll
        for (int k = 0; k < NW; ++k) {
          uv[k] += a[k];
          vv[k] += b[k];
        }
      }
#pragma unroll
      for (int k = 0; k < NW; ++k)
        if (f >= win[k]) tile[k][lx][al] = soap_epilogue(uv[k], uu, vv[k], kind, power);
    }
  }
  __syncthreads();
  for (int i = 0; i < 4; ++i) {
    const int fl = ly + 8 * i;
    const long long a = a0 + lx, f = f0 + fl;
    if (a < A && f < F) {
#pragma unroll
      for (int k = 0; k < NW; ++k)
        if (f >= win[k]) t[off[k] + (size_t)(f - win[k]) * A + a] = tile[k][fl][lx];
    }
  }
}

// dt[a][r] = t[r+1][a] - t[r][a]   (numpy.diff(t.T, axis=-1))
__global__ void __launch_bounds__(256)
    diff_transposed_kernel(const double *__restrict__ t, double *__restrict__ dt, long long rows, long long A) {
  __shared__ double tile[33][33];
  const long long r0 = (long long)blockIdx.x * 32, a0 = (long long)blockIdx.y * 32;
  const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
  for (int rl = ly; rl < 33; rl += 8) {
    const long long r = r0 + rl, a = a0 + lx;
    if (r < rows && a < A) tile[rl][lx] = t[(size_t)r * A + a];
  }
  __syncthreads();
  for (int al = ly; al < 32; al += 8) {
    const long long a = a0 + al, r = r0 + lx;
    if (a < A && r + 1 < rows) dt[(size_t)a * (rows - 1) + r] = tile[lx + 1][al] - tile[lx][al];
  }
}

}  // namespace soapb
#include "timelag_rb.cuh"
#include "timelag_pair.cuh"
#include "timelag_chain.cuh"
namespace soapb {

// The register-blocked kernel is OPT-IN (SOAP_TL_RB=1): on B200 it is slower than timelag_kernel (0.43 against
// 0.79 of the HBM peak for the 4-window fp64 sweep, profiles/r02_timelag_rb_ncu.txt).  It does cut the shared-memory
// wavefronts as designed (pipe 35 % busy), but 7 frames per lane means 112-frame blocks, i.e. only 4 consumer
// warps = ONE per scheduler, and a single warp issues an instruction every ~4.7 cycles on this SM (fixed-latency
// "wait" stalls between dependent issues that a second warp would hide).  A second warp per scheduler doubles the
// per-block hand-over of partial sums (35 values per lane for 3.5 units of work).  Kept, with its parity tests, as
// the measured record of that design; it applies to TMA-able rows, cosine-type kinds, windows <= 56 frames with
// at most one of them beyond the 10-frame register history.
struct RbChoice {
  bool use;
  int ns, nl, hist;
  int w[4];     // windows sorted ascending
  int perm[4];  // w[k] = windows[perm[k]]
};
static RbChoice rb_choose(int64_t F, int64_t A, int D, int es, const int *windows, int nw, int kind, bool bulk_ok) {
  RbChoice ch;
  memset(&ch, 0, sizeof(ch));
  const int mode = env_int("SOAP_TL_RB", -1);
  if (mode != 1 || !bulk_ok || kind == SOAP_KIND_EUCLID || kind == SOAP_KIND_EUCLID_NORMALIZED || nw < 1 || nw > 4) return ch;
  for (int k = 0; k < nw; ++k) ch.perm[k] = k;
  for (int i = 1; i < nw; ++i)
    for (int j = i; j > 0 && windows[ch.perm[j]] < windows[ch.perm[j - 1]]; --j) {
      const int t = ch.perm[j];
      ch.perm[j] = ch.perm[j - 1];
      ch.perm[j - 1] = t;
    }
  for (int k = 0; k < nw; ++k) {
    ch.w[k] = windows[ch.perm[k]];
    if (ch.w[k] <= kRbHist) ++ch.ns, ch.hist = ch.w[k];
    else if (ch.w[k] <= kRbMaxLag) ++ch.nl;
    else return ch;
  }
  if (ch.nl > 1) return ch;
  if ((D * es) % 16) return ch;
  if ((int64_t)A * D * es / 8 >= 2147483647LL) return ch;
  ch.use = true;
  return ch;
}

template <typename T, int NS, int NL>
static int launch_timelag_rb(const void *X, const void *mult, double *scratch, int64_t F, int64_t A, int D, const RbPlan &pl,
                             const RbChoice &ch, cudaStream_t s) {
  auto kern = timelag_rb_kernel<T, NS, NL>;
  SOAP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  long long grid = num_sms();
  const int force_grid = env_int("SOAP_TL_GRID", 0);
  if (force_grid > 0) grid = force_grid;
  if (grid > A) grid = A;
  SOAP_REQUIRE(grid <= tl_max_grid(), "soap_timelag: grid %lld exceeds the accumulator area", grid);
  const int64_t nblk = (F + kRbB - 1) / kRbB;
  SOAP_REQUIRE(((A + grid - 1) / grid) * pl.pieces * nblk < 2147483647LL, "soap_timelag: too many blocks per CTA");
  // scratch: totals sums[a][q][f] | per-CTA accumulators [grid][q][frames padded to blocks] | shifted multiplicities
  double *cta_acc = scratch + (size_t)A * pl.Q * (size_t)F;
  const size_t padded = (size_t)nblk * kRbB;
  T *table = reinterpret_cast<T *>((reinterpret_cast<uintptr_t>(cta_acc + (size_t)tl_max_grid() * pl.Q * padded) + 127) & ~(uintptr_t)127);
  const int units_total = pl.pieces * pl.c * 8;
  rb_mult_setup_kernel<T><<<32, 256, 0, s>>>(static_cast<const T *>(mult), table, D, units_total);
  int rc = check_launch("rb_mult_setup_kernel");
  if (rc) return rc;
  // X as [F][A*D] 8-byte words; one map per box width (1..4 lines of 16 words), 56 frames per box, no L2 promotion
  CUtensorMap tm[4];
  memset(tm, 0, sizeof(tm));
  const int64_t row_words = (int64_t)A * D * (int64_t)sizeof(T) / 8;
  for (int st = 0; st < pl.st.n; ++st) {
    const int k = pl.st.k[st];
    rc = make_tensor_map_2d(&tm[k - 1], X, CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, F, row_words, 16 * k, kRbBox,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE);
    if (rc) return rc;
  }
  const int4 win4 = make_int4(ch.w[0], ch.w[1], ch.w[2], ch.w[3]);
  ProfScope prof(SOAP_PROF_TIMELAG, s);
  kern<<<(unsigned)grid, kRbThreads, pl.smem, s>>>(table, scratch, cta_acc, (long long)F, (long long)A, pl.n_units, pl.c,
                                                   pl.pieces, (int)pl.mult_elems, win4, ch.hist, env_int("SOAP_RB_DIAG", 0), pl.st, tm[0], tm[1], tm[2], tm[3]);
  g_tl_last_kernel = "timelag_rb_kernel";
  return check_launch("timelag_rb_kernel");
}

template <typename T>
static int dispatch_timelag_rb(const void *X, const void *mult, double *scratch, int64_t F, int64_t A, int D, const RbPlan &pl,
                               const RbChoice &ch, cudaStream_t s) {
#define RB_CASE(NS, NL) \
  if (ch.ns == NS && ch.nl == NL) return launch_timelag_rb<T, NS, NL>(X, mult, scratch, F, A, D, pl, ch, s);
  RB_CASE(1, 0) RB_CASE(2, 0) RB_CASE(3, 0) RB_CASE(4, 0) RB_CASE(0, 1) RB_CASE(1, 1) RB_CASE(2, 1) RB_CASE(3, 1)
#undef RB_CASE
  return set_error(SOAP_EINVAL, "soap_timelag: no register-blocked variant for %d + %d windows", ch.ns, ch.nl);
}

// inv[a][f] = 1 / sqrt(sum_e m_e x[f][a][e]^2), 1 for a zero row (normalizeArray, utils.py:149-151): the norms of the
// EXPANDED vectors from the compressed rows; one warp per row, 128-bit loads when the rows allow
template <typename T>
__global__ void __launch_bounds__(256)
    row_inv_norm_kernel(const T *__restrict__ X, const T *__restrict__ mult, double *__restrict__ inv, long long F, long long A,
                        int D, int vec_ok) {
  constexpr int UE = 16 / (int)sizeof(T);
  const int lane = threadIdx.x & 31;
  const long long nrows = F * A, nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < nrows; r += nwarps) {
    const T *x = X + (size_t)r * D;
    double acc = 0.0;
    if (vec_ok) {
      for (int u = lane; u < D / UE; u += 32) {
        const int4 raw = ldg_stream16(x + u * UE);
        const T *v = reinterpret_cast<const T *>(&raw);
#pragma unroll
        for (int e = 0; e < UE; ++e) {
          const double xv = (double)v[e], mv = mult ? (double)mult[u * UE + e] : 1.0;
          acc = fma(xv * mv, xv, acc);
        }
      }
    } else {
      for (int e = lane; e < D; e += 32) {
        const double xv = (double)x[e], mv = mult ? (double)mult[e] : 1.0;
        acc = fma(xv * mv, xv, acc);
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) {
      const long long f = r / A, a = r - f * A;
      inv[(size_t)a * F + f] = acc == 0.0 ? 1.0 : 1.0 / sqrt(acc);
    }
  }
}

// shifted rows: table[lead][i] = multiplicity of element i - lead of the row (ones when the caller passes none),
// zero outside the row -- what masks the foreign elements of the first / last unit of an atom's window
template <typename T>
__global__ void shift_mult_kernel(const T *__restrict__ mult, T *__restrict__ out, int D, int pitch, int n_leads) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_leads * pitch; i += gridDim.x * blockDim.x) {
    const int lead = i / pitch, e = i - lead * pitch - lead;
    out[i] = (e >= 0 && e < D) ? (mult ? mult[e] : (T)1) : (T)0;
  }
}

template <typename T, int NW, bool HM>
static int launch_timelag(const void *X, const void *mult, double *scratch, int64_t F, int64_t A, int D,
                          const TimelagPlan &pl, const int *w, int kind, int bulk_ok, int mult_pitch, const double *inv_norm,
                          cudaStream_t s) {
  const int4 win4 = make_int4(w[0], NW > 1 ? w[1] : 0, NW > 2 ? w[2] : 0, NW > 3 ? w[3] : 0);
  auto kern = (kind == SOAP_KIND_EUCLID) ? timelag_kernel<T, NW, 1, HM, kTlWarps, 4>
              : (kind == SOAP_KIND_EUCLID_NORMALIZED) ? timelag_kernel<T, NW, 2, HM, kTlWarps, 4>
                                                      : timelag_kernel<T, NW, 0, HM, kTlWarps, 4>;
  int cw = kTlWarps;
  if constexpr (NW == 4 && sizeof(T) == 8 && HM) {  // launch-shape experiments (scripts/tl_sweep.py)
    if (kind != SOAP_KIND_EUCLID && kind != SOAP_KIND_EUCLID_NORMALIZED && pl.cw != kTlWarps) {
      cw = pl.cw;
#define TL_VAR(W, B) if (cw == W) kern = timelag_kernel<T, NW, 0, HM, W, B>;
      TL_VAR(4, 4) TL_VAR(6, 4) TL_VAR(10, 4) TL_VAR(12, 4)
#undef TL_VAR
    }
  }
  SOAP_REQUIRE(cw == pl.cw, "soap_timelag: SOAP_TL_WARPS=%d is only built for the 4-window f64 sweep", pl.cw);
  const int threads = (cw + 2) * 32;
  SOAP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  // persistent grid: as many CTAs as fit on the device at once (one per SM with a full-size ring)
  int dev = 0, sms = 0, per_sm = 0;
  SOAP_CUDA(cudaGetDevice(&dev));
  SOAP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  SOAP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, pl.smem));
  SOAP_REQUIRE(per_sm >= 1, "soap_timelag: kernel does not fit on an SM (%zu bytes of shared memory)", pl.smem);
  long long grid = (long long)(sms - (num_sms() - persistent_sms())) * per_sm;  // soap_reserve_sms()
  if (grid < 1) grid = 1;
  const int force_grid = env_int("SOAP_TL_GRID", 0);
  if (force_grid > 0) grid = force_grid;
  const bool atom_major = tl_atom_major(A);
  const long long n_items = (long long)A * pl.split;
  if (grid > (atom_major ? (long long)A : n_items)) grid = atom_major ? (long long)A : n_items;
  SOAP_REQUIRE(!atom_major || grid <= tl_max_grid(), "soap_timelag: grid %lld exceeds the accumulator area", grid);
  SOAP_REQUIRE((n_items + grid - 1) / grid * ((F + kBlk - 1) / kBlk) + (long long)pl.split * ((F + kBlk - 1) / kBlk) < 2147483647LL,
               "soap_timelag: too many blocks per CTA (%lld atoms x %d pieces)", (long long)A, pl.split);
  double *cta_acc = atom_major ? scratch + (size_t)A * pl.Q * (size_t)F : nullptr;
  // X viewed as a 2-D tensor [F][A*D] of 8-byte words; box = 32 frames x one piece (stride bytes).  No L2
  // promotion: the neighbouring pieces of a row are fetched by the same CTA one item later, long after the
  // promoted sectors have left L2 (measured: DRAM reads 1.12x -> 1.08x the algorithmic bytes, (1,) 0.965 -> 0.988)
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  if (bulk_ok & 1) {
    const int64_t row_words = (int64_t)A * D * (int64_t)sizeof(T) / 8;
    SOAP_REQUIRE(row_words < 2147483647LL, "soap_timelag: a frame of %lld bytes exceeds the tensor-map coordinate range",
                 (long long)row_words * 8);
    const int rc = make_tensor_map_2d(&tmap, X, CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, F, row_words, pl.stride / 8, kBlk,
                                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE);
    if (rc) return rc;
  }
  ProfScope prof(SOAP_PROF_TIMELAG, s);
  kern<<<(unsigned)grid, threads, pl.smem, s>>>(static_cast<const T *>(X), static_cast<const T *>(mult), scratch, cta_acc,
                                                inv_norm, (long long)F, (long long)A, D, pl.split, pl.upp, pl.stride, pl.R,
                                                (pl.max_lag + kBlk - 1) / kBlk, pl.mbuf, win4, bulk_ok, mult_pitch, tmap);
  g_tl_last_kernel = "timelag_kernel";
  return check_launch("timelag_kernel");
}

// ---- two frames per lane (timelag_pair.cuh) --------------------------------------------------------------------------
struct PairChoice {
  bool use;
  int w[4], perm[4];  // w[k] = windows[perm[k]], window 1 first
};
static PairChoice pair_choose(int64_t F, int nw, const int *windows, int kind, int bulk_ok, int es) {
  PairChoice ch;
  memset(&ch, 0, sizeof(ch));
  // SOAP_TL_PAIR: 1 = use it whenever it applies, 0 = never.  Default: float32 data only -- interleaved A/B on
  // 1000 x 4096 x 1224, windows 1,5,10,50 (profiles/r02_timelag_pair_ab.txt): fp32 4.01 / 4.26 / 4.38 ms against
  // 4.23 / 4.47 / 4.66 (5-6 % faster, every round), fp64 8.90-8.96 against 8.79-8.97 (no gain: at 162-168 registers
  // the fp64 variant cannot keep two units of loads in flight; the microbenchmark of the bare loops promised 12 %)
  const int mode = env_int("SOAP_TL_PAIR", kPairDefault);
  const bool on = mode == 1 || (mode < 0 && es == 4 && env_int("SOAP_TL_RB", -1) != 1);  // (an explicit SOAP_TL_RB=1 wins)
  if (!on || bulk_ok != 1 || (F & 1) || nw < 3 || nw > 4 || kind == SOAP_KIND_EUCLID || kind == SOAP_KIND_EUCLID_NORMALIZED) return ch;
  int one = -1;
  for (int k = 0; k < nw; ++k) {
    if (windows[k] > kPrBlk) return ch;
    if (windows[k] == 1) one = k;
  }
  if (one < 0) return ch;
  int n = 0;
  ch.perm[n++] = one;
  for (int k = 0; k < nw; ++k)
    if (k != one) ch.perm[n++] = k;
  for (int k = 0; k < nw; ++k) ch.w[k] = windows[ch.perm[k]];
  ch.use = true;
  return ch;
}

template <typename T, int NW, bool HM>
static int launch_timelag_pair(const void *X, const void *mult, double *scratch, int64_t F, int64_t A, int D, const PairPlan &pl,
                               const PairChoice &ch, cudaStream_t s) {
  auto kern = timelag_pair_kernel<T, NW, HM, true, 2>;
  if constexpr (NW == 4 && HM) {  // load-batch experiments (scripts/tl_sweep.py): fp32 1 / 2 / 3 units = 4.90 / 4.14 / 4.74 ms
    const int ub = env_int("SOAP_TL_PAIR_UB", 2);
    if (ub == 1) kern = timelag_pair_kernel<T, NW, HM, true, 1>;
    if (ub == 3) kern = timelag_pair_kernel<T, NW, HM, true, 3>;
  }
  SOAP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  const int threads = (kPrWarps + 2) * 32;
  int per_sm = 0;
  SOAP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, pl.smem));
  SOAP_REQUIRE(per_sm >= 1, "soap_timelag: the pair kernel does not fit on an SM (%zu bytes of shared memory)", pl.smem);
  long long grid = (long long)persistent_sms() * per_sm;
  if (grid < 1) grid = 1;
  const int force_grid = env_int("SOAP_TL_GRID", 0);
  if (force_grid > 0) grid = force_grid;
  const bool atom_major = tl_atom_major(A);
  const long long n_items = (long long)A * pl.split;
  if (grid > (atom_major ? (long long)A : n_items)) grid = atom_major ? (long long)A : n_items;
  SOAP_REQUIRE(!atom_major || grid <= tl_max_grid(), "soap_timelag: grid %lld exceeds the accumulator area", grid);
  const int64_t nblk = (F + kPrBlk - 1) / kPrBlk;
  SOAP_REQUIRE((n_items + grid - 1) / grid * nblk + (long long)pl.split * nblk < 2147483647LL, "soap_timelag: too many blocks per CTA");
  double *cta_acc = atom_major ? scratch + (size_t)A * pl.Q * (size_t)F : nullptr;
  const int64_t row_words = (int64_t)A * D * (int64_t)sizeof(T) / 8;
  SOAP_REQUIRE(row_words < 2147483647LL, "soap_timelag: a frame of %lld bytes exceeds the tensor-map coordinate range", (long long)row_words * 8);
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  int rc = make_tensor_map_pairs(&tmap, X, F / 2, row_words, pl.stride / 8);
  if (rc) return rc;
  const int4 win4 = make_int4(ch.w[0], ch.w[1], ch.w[2], NW > 3 ? ch.w[3] : 0);
  ProfScope prof(SOAP_PROF_TIMELAG, s);
  kern<<<(unsigned)grid, threads, pl.smem, s>>>(static_cast<const T *>(mult), scratch, cta_acc, (long long)F, (long long)A, D, pl.split, pl.upp,
                                                pl.stride, pl.mbuf, win4, tmap);
  g_tl_last_kernel = "timelag_pair_kernel";
  return check_launch("timelag_pair_kernel");
}


// ---- lane = units, lag history in registers (timelag_chain.cuh) ------------------------------------------------------
// SOAP_TL_CHAIN_U = 3 | 4 units per lane, SOAP_TL_CHAIN_SR = 1 | 0: partial sums through shared memory | by shuffles
// (ring window?, chain windows / 5): (w0, 5, 50), (w0, 10, 50), (w0, 5, 10, 25), (w0, 5, 10, 20), (5, 10, 50), (5, 10, 25)
#define CH_MORE_PATTERNS CH_MORE(1, 1, 10, 0) CH_MORE(1, 2, 10, 0) CH_MORE(1, 1, 2, 5) CH_MORE(1, 1, 2, 4) CH_MORE(0, 1, 2, 10) CH_MORE(0, 1, 2, 5)
struct ChainChoice {
  bool use;
  int nw, ns, w0, c[3];  // ring window (ns = 1) and chain windows c[k] * 5, ascending
  int w[4], perm[4];     // w[k] = windows[perm[k]] in the kernel's order: ring window first, then the chain windows
};
static ChainChoice chain_choose(int64_t F, int64_t A, int nw, const int *windows, int kind, int bulk_ok, int es) {
  ChainChoice ch;
  memset(&ch, 0, sizeof(ch));
  // SOAP_TL_CHAIN: 1 = use it whenever it applies, 0 = never.  Default: float64 data only -- bench.py, 1000 x 12 500 x
  // 1224 fp64, windows 1,5,10,50 (profiles/r03_bench_chain_ab.txt): 21.8 ms = 0.86 of the HBM peak sustained (0.91
  // timed alone) against 25.3 ms = 0.74 (0.78) for timelag_kernel; float32 4.5 ms against 4.0 ms for the pair kernel
  // (1000 x 4096: the fp32 variant issues twice the instructions per byte and five warps cannot hide them)
  const int mode = env_int("SOAP_TL_CHAIN", kChainDefault);
  const bool on = mode == 1 || (mode < 0 && es == 8 && env_int("SOAP_TL_RB", -1) != 1 && env_int("SOAP_TL_PAIR", -1) != 1);
  if (!on || bulk_ok != 1 || nw < 2 || nw > 4 || kind == SOAP_KIND_EUCLID || kind == SOAP_KIND_EUCLID_NORMALIZED) return ch;
  // a CTA owns whole atoms here: with fewer than 4 atoms per SM the (atom, piece) items of timelag_kernel keep more
  // SMs busy -- unless the caller insists (bench.py recomputes 16 atoms of a neighbour's shard BITWISE)
  if (mode != 1 && !tl_atom_major(A)) {
    // ... or when whole atoms fill the SMs evenly enough (148 atoms: one each, 0.94 of the HBM peak against 0.75; 400: 0.89 against 0.73)
    const long long sms = persistent_sms() > 0 ? persistent_sms() : 1;
    const double busy = (double)A / (double)(((A + sms - 1) / sms) * sms);
    if (busy < 0.72) return ch;  // (break-even measured at 200 atoms = 0.68: profiles/r03_timelag_chain_small_a.txt)
  }
  int ring = -1, nc = 0, ci[3] = {0, 0, 0};
  for (int k = 0; k < nw; ++k) {
    const int w = windows[k];
    if (w % kChS == 0 && w <= kChS * kChH) {
      if (nc == 3) return ch;
      ci[nc++] = k;
    } else {
      if (ring >= 0 || w > kChB) return ch;
      ring = k;
    }
  }
  for (int i = 1; i < nc; ++i)
    for (int j = i; j > 0 && windows[ci[j]] < windows[ci[j - 1]]; --j) {
      const int t = ci[j];
      ci[j] = ci[j - 1];
      ci[j - 1] = t;
    }
  for (int i = 1; i < nc; ++i)
    if (windows[ci[i]] == windows[ci[i - 1]]) return ch;  // (a repeated window takes the general kernel)
  int n = 0;
  if (ring >= 0) ch.perm[n++] = ring, ch.ns = 1, ch.w0 = windows[ring];
  for (int i = 0; i < nc; ++i) ch.perm[n++] = ci[i], ch.c[i] = windows[ci[i]] / kChS;
  for (int k = 0; k < nw; ++k) ch.w[k] = windows[ch.perm[k]];
  ch.nw = nw;
  // instantiated patterns (launch_chain_pattern)
  const bool p4 = ch.ns == 1 && ch.c[0] == 1 && ch.c[1] == 2 && ch.c[2] == 10;  // (w0, 5, 10, 50)
  const bool p3 = ch.ns == 1 && ch.c[0] == 1 && ch.c[1] == 2 && ch.c[2] == 0;   // (w0, 5, 10)
  // further patterns, float64 and the default variant only (every pattern is an instantiation of its own)
  bool more = false;
  if (es == 8 && env_int("SOAP_TL_CHAIN_SR", 1) == 1 && env_int("SOAP_TL_CHAIN_U", 4) == 4 && nw >= 3) {
#define CH_MORE(NS, C1, C2, C3) more = more || (ch.ns == NS && ch.c[0] == C1 && ch.c[1] == C2 && ch.c[2] == C3);
    CH_MORE_PATTERNS
#undef CH_MORE
  }
  ch.use = p4 || p3 || more;
  return ch;
}

template <typename T, int U, int SR, int NS, int C1, int C2, int C3, bool HM>
static int launch_timelag_chain(const void *X, const void *mult, double *scratch, int64_t F, int64_t A, int D, const ChainPlan &pl,
                                const ChainChoice &ch, int ov, int halo, cudaStream_t s) {
  // SR: 0 = shuffle butterfly, 1 = sums through shared memory, 2 + d = the same, loads pipelined d columns ahead
  auto kern = timelag_chain_kernel<T, U, NS, C1, C2, C3, HM, (SR > 0), (SR >= 2 ? SR - 2 : 0)>;
  SOAP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  int per_sm = 0;
  SOAP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kChThreads, pl.smem));
  SOAP_REQUIRE(per_sm >= 1, "soap_timelag: the chain kernel does not fit on an SM (%zu bytes of shared memory)", pl.smem);
  long long grid = (long long)persistent_sms() * per_sm;
  if (grid < 1) grid = 1;
  const int force_grid = env_int("SOAP_TL_GRID", 0);
  if (force_grid > 0) grid = force_grid;
  if (grid > (long long)A) grid = (long long)A;
  SOAP_REQUIRE(grid <= tl_max_grid(), "soap_timelag: grid %lld exceeds the accumulator area", grid);
  const int64_t nblk = (F + kChB - 1) / kChB;
  SOAP_REQUIRE(((long long)A * pl.split + grid - 1) / grid * nblk + (long long)pl.split * nblk < 2147483647LL, "soap_timelag: too many blocks per CTA");
  double *cta_acc = scratch + (size_t)A * pl.Q * (size_t)F;
  const int64_t row_words = (int64_t)A * D * (int64_t)sizeof(T) / 8;
  SOAP_REQUIRE(row_words < 2147483647LL, "soap_timelag: a frame of %lld bytes exceeds the tensor-map coordinate range", (long long)row_words * 8);
  CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  CUtensorMap tmap_last, tmap_first, tmap_last4;
  memset(&tmap_last, 0, sizeof(tmap_last));
  memset(&tmap_first, 0, sizeof(tmap_first));
  memset(&tmap_last4, 0, sizeof(tmap_last4));
  // SOAP_TL_CHAIN_ALIGN=1: rows of 8 k + 4 units (every other atom starts 64 bytes into a 128-byte line): that atom's
  // pieces are cut 4 units early, so that every boundary inside a row falls on a line (ch_item)
  const int n_units = D / (16 / (int)sizeof(T));
  const int align = (env_int("SOAP_TL_CHAIN_ALIGN", kChainAlignDefault) != 0 && pl.split > 1 && (n_units & 7) == 4 &&
                     (reinterpret_cast<uintptr_t>(X) & 127) == 0 && pl.last_units + 4 <= pl.upp)
                        ? 1 : 0;
  int rc = make_tensor_map_2d(&tmap, X, CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, F, row_words, pl.pitch / 8, kChB + ov, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_NONE);
  if (rc) return rc;
  rc = make_tensor_map_2d(&tmap_last, X, CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, F, row_words, pl.last_units * 2, kChB + ov, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_NONE);
  if (rc) return rc;
  if (align) {
    rc = make_tensor_map_2d(&tmap_first, X, CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, F, row_words, (pl.upp - 4) * 2, kChB + ov, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_NONE);
    if (rc) return rc;
    rc = make_tensor_map_2d(&tmap_last4, X, CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, F, row_words, (pl.last_units + 4) * 2, kChB + ov, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_NONE);
    if (rc) return rc;
  }
  const int4 pitch4 = make_int4(pl.pitch, (pl.upp - 4) * 16, pl.last_units * 16, (pl.last_units + 4) * 16);
  ProfScope prof(SOAP_PROF_TIMELAG, s);
  kern<<<(unsigned)grid, kChThreads, pl.smem, s>>>(static_cast<const T *>(mult), scratch, cta_acc, (long long)F, (long long)A, D, pl.split, pl.upp,
                                                   pitch4, (int)pl.blk_bytes, pl.nbuf, ch.w0, ov, env_int("SOAP_TL_CHAIN_RED", 2) == 1 ? 1 : 2, halo, align, tmap,
                                                   tmap_first, tmap_last, tmap_last4);
  g_tl_last_kernel = "timelag_chain_kernel";
  return check_launch("timelag_chain_kernel");
}

template <typename T, int U, int SR>
static int launch_chain_pattern(const void *X, const void *mult, double *scratch, int64_t F, int64_t A, int D, const ChainPlan &pl,
                                const ChainChoice &ch, int ov, int halo, cudaStream_t s) {
#define CH_CASE(NS, C1, C2, C3)                                                                                   \
  if (ch.ns == NS && ch.c[0] == C1 && ch.c[1] == C2 && ch.c[2] == C3)                                             \
    return mult ? launch_timelag_chain<T, U, SR, NS, C1, C2, C3, true>(X, mult, scratch, F, A, D, pl, ch, ov, halo, s)  \
                : launch_timelag_chain<T, U, SR, NS, C1, C2, C3, false>(X, mult, scratch, F, A, D, pl, ch, ov, halo, s);
  CH_CASE(1, 1, 2, 10)
  CH_CASE(1, 1, 2, 0)
  if constexpr (std::is_same<T, double>::value && U == 4 && SR == 1) {
#define CH_MORE(NS, C1, C2, C3) CH_CASE(NS, C1, C2, C3)
    CH_MORE_PATTERNS
#undef CH_MORE
  }
#undef CH_CASE
  return set_error(SOAP_EINVAL, "soap_timelag: no chain variant for this window pattern");
}

template <int NW>
static int launch_finalize(const double *scratch, double *t, const int64_t *off, int64_t F, int64_t A,
                           int split, const int *w, int kind, int power, cudaStream_t s) {
  const int4 win4 = make_int4(w[0], NW > 1 ? w[1] : 0, NW > 2 ? w[2] : 0, NW > 3 ? w[3] : 0);
  dim3 grid((unsigned)((F + 31) / 32), (unsigned)((A + 31) / 32));
  SOAP_REQUIRE(grid.y <= 65535, "soap_timelag: more than 2M atoms per call are not supported");
  timelag_finalize_kernel<NW><<<grid, 256, 0, s>>>(scratch, t, off[0], NW > 1 ? off[1] : 0, NW > 2 ? off[2] : 0,
                                                  NW > 3 ? off[3] : 0, (long long)F, (long long)A, split, win4, kind, power);
  return check_launch("timelag_finalize_kernel");
}

}  // namespace soapb

using namespace soapb;

// room at the end of the scratch area for the zero-padded multiplicities (rows that are not whole units)
static size_t tl_pad_bytes(int dim, int es) { return 4 * ((size_t)dim + 32) * es + 256; }
// ... and, before it, for the reciprocal row norms of SOAP_KIND_EUCLID_NORMALIZED
static size_t tl_inv_bytes(int64_t F, int64_t A) { return (size_t)F * A * sizeof(double) + 256; }

static int check_windows(int64_t F, const int *w, int nw, const char *who) {
  SOAP_REQUIRE(w != nullptr && nw >= 1 && nw <= SOAP_MAX_WINDOWS, "%s: 1..%d windows per call", who, SOAP_MAX_WINDOWS);
  for (int k = 0; k < nw; ++k)
    SOAP_REQUIRE(w[k] >= 1 && w[k] < F, "%s: window %d must be in [1, n_frames)", who, w[k]);
  return SOAP_OK;
}

extern "C" const char *soap_timelag_last_kernel(void) { return g_tl_last_kernel; }

extern "C" size_t soap_timelag_scratch_bytes(int64_t n_frames, int64_t n_atoms, int dim, int dtype,
                                             const int *windows_host, int n_windows) {
  if (n_frames <= 0 || n_atoms <= 0 || dim <= 0 || !windows_host || n_windows < 1 || n_windows > SOAP_MAX_WINDOWS)
    return 0;
  const int es = dtype == SOAP_F64 ? 8 : 4;
  const int Q = n_windows + 1;
  // register-blocked kernel (whatever the caller's pointers and kind turn out to be, the area covers it): totals
  // sums[a][q][f] + one accumulator per resident CTA + the 8 shifted multiplicity tables
  size_t rb_bytes = 0;
  RbPlan rp;
  if (rb_make_plan(dim, es, n_windows, rp)) {
    const size_t padded = (size_t)((n_frames + kRbB - 1) / kRbB) * kRbB;
    rb_bytes = ((size_t)n_atoms * Q * (size_t)n_frames + (size_t)tl_max_grid() * Q * padded) * sizeof(double) +
               8 * rp.mult_elems * es + 256;
  }
  TimelagPlan pl;
  const int plan_dim = ((size_t)dim * es) % 16 != 0 ? dim + 16 / es - 1 : dim;  // shifted rows: the widest window
  if (!make_plan(n_frames, plan_dim, es, windows_host, n_windows, pl))
    return rb_bytes + tl_pad_bytes(dim, es) + tl_inv_bytes(n_frames, n_atoms);
  int pair_split = 0;  // the pair kernel cuts rows into more pieces (timelag_pair.cuh)
  {
    PairPlan pp;
    if (pr_make_plan(n_frames, plan_dim, es, n_windows, pp)) pair_split = pp.split;
  }
  // totals sums[a][q][f] + one accumulator [q][frames padded to a block: 32, or 64 for the pair kernel] per resident CTA
  // (the chain kernel's 25-frame blocks and up to 5-piece rows fit the same area: F + 24 <= padded + 24)
  const size_t padded = (size_t)((n_frames + kPrBlk - 1) / kPrBlk) * kPrBlk + kChB;
  size_t bytes = tl_atom_major(n_atoms)
                     ? ((size_t)n_atoms * pl.Q * (size_t)n_frames + (size_t)tl_max_grid() * pl.Q * padded) * sizeof(double)
                     : (size_t)n_atoms * (size_t)(pl.split > pair_split ? pl.split : pair_split) * pl.Q * (size_t)n_frames * sizeof(double);
  // the chain kernel forced on for a few atoms (SOAP_TL_CHAIN=1): totals + one accumulator per CTA (at most one per atom)
  const size_t forced = ((size_t)n_atoms * pl.Q * (size_t)n_frames +
                         (size_t)(n_atoms < tl_max_grid() ? n_atoms : tl_max_grid()) * pl.Q * padded) * sizeof(double);
  if (forced > bytes) bytes = forced;
  return (bytes > rb_bytes ? bytes : rb_bytes) + tl_pad_bytes(dim, es) + tl_inv_bytes(n_frames, n_atoms);
}

extern "C" int soap_timelag(const void *X, const void *mult, double *t, const int64_t *t_offsets_host,
                            int64_t n_frames, int64_t n_atoms, int dim, const int *windows_host,
                            int n_windows, int kind, int power, int dtype, void *scratch, void *stream) {
  SOAP_REQUIRE(n_frames > 0 && n_atoms > 0 && dim > 0, "soap_timelag: bad shape");
  SOAP_REQUIRE(X && t && t_offsets_host && scratch, "soap_timelag: null buffer");
  SOAP_REQUIRE(dtype == SOAP_F32 || dtype == SOAP_F64, "soap_timelag: dtype %d", dtype);
  SOAP_REQUIRE(kind >= SOAP_KIND_SIMPLE && kind <= SOAP_KIND_EUCLID_NORMALIZED, "soap_timelag: kind %d", kind);
  int rc = check_windows(n_frames, windows_host, n_windows, "soap_timelag");
  if (rc) return rc;
  const int es = dtype == SOAP_F64 ? 8 : 4;
  // TMA path: 16-byte aligned base pointers and frames of whole 16-byte units (the tensor map counts 8-byte words);
  // rows need not be whole units: a window then starts at the unit boundary below its row and the foreign elements
  // of its first / last unit are masked (tl_item)
  const bool whole_units = ((size_t)dim * es) % 16 == 0;
  int bulk_ok = ((reinterpret_cast<uintptr_t>(X) & 15) == 0) && (((size_t)n_atoms * dim * es) % 16 == 0) &&
                ((reinterpret_cast<uintptr_t>(mult) & 15) == 0);
  auto s = static_cast<cudaStream_t>(stream);
  auto sc = static_cast<double *>(scratch);
  const size_t scratch_total = soap_timelag_scratch_bytes(n_frames, n_atoms, dim, dtype, windows_host, n_windows);
  const double *inv_norm = nullptr;
  if (kind == SOAP_KIND_EUCLID_NORMALIZED) {
    // timeSOAPsimple on normalizeArray(fillSOAPVectorFromdscribe(x)) without materialising it: one pass for the
    // norms of the expanded vectors (from the compressed rows and the multiplicities), then |x/|x| - y/|y||
    double *inv = reinterpret_cast<double *>(
        (reinterpret_cast<uintptr_t>(static_cast<char *>(scratch) + scratch_total - tl_pad_bytes(dim, es) - tl_inv_bytes(n_frames, n_atoms)) + 127) & ~(uintptr_t)127);
    const int vec_ok = ((size_t)dim * es) % 16 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(mult) & 15) == 0;
    ProfScope prof(SOAP_PROF_EXPAND, s);
    if (dtype == SOAP_F64)
      row_inv_norm_kernel<double><<<num_sms() * 8, 256, 0, s>>>(static_cast<const double *>(X), static_cast<const double *>(mult), inv, n_frames, n_atoms, dim, vec_ok);
    else
      row_inv_norm_kernel<float><<<num_sms() * 8, 256, 0, s>>>(static_cast<const float *>(X), static_cast<const float *>(mult), inv, n_frames, n_atoms, dim, vec_ok);
    rc = check_launch("row_inv_norm_kernel");
    if (rc) return rc;
    inv_norm = inv;
  }
  int mult_pitch = 0, plan_dim = dim;
  if (bulk_ok && !whole_units) {
    // rows that are not whole units: an atom's window starts up to ue - 1 elements early (tl_item); the plan covers
    // the widest window, the multiplicities become a [ue][pitch] table that is zero on the foreign elements
    const int ue = 16 / es;
    plan_dim = dim + ue - 1;
    mult_pitch = (plan_dim + ue - 1) / ue * ue + 4 * ue;
    void *table = reinterpret_cast<void *>((reinterpret_cast<uintptr_t>(static_cast<char *>(scratch) + scratch_total - tl_pad_bytes(dim, es)) + 127) & ~(uintptr_t)127);
    if (dtype == SOAP_F64)
      shift_mult_kernel<double><<<8, 256, 0, s>>>(static_cast<const double *>(mult), static_cast<double *>(table), dim, mult_pitch, ue);
    else
      shift_mult_kernel<float><<<8, 256, 0, s>>>(static_cast<const float *>(mult), static_cast<float *>(table), dim, mult_pitch, ue);
    rc = check_launch("shift_mult_kernel");
    if (rc) return rc;
    mult = table;
    bulk_ok |= 32;
  }
  {
    const ChainChoice cc = chain_choose(n_frames, n_atoms, n_windows, windows_host, kind, bulk_ok, es);
    ChainPlan cp;
    const int chU = env_int("SOAP_TL_CHAIN_U", 4);
    const int chSR = env_int("SOAP_TL_CHAIN_SR", 1);
    // SOAP_TL_CHAIN_OV=0: no overlapping boxes (the previous block is held back for the ring-window partners instead)
    // (window 2 with full-width pieces: three buffers of 27 rows do not fit beside the partial sums -> no overlap)
    int chOV = (cc.ns && cc.w0 <= 2 && env_int("SOAP_TL_CHAIN_OV", 1) != 0) ? cc.w0 : 0;
    bool ch_ok = cc.use && whole_units && (chU == 3 || chU == 4);
    // SOAP_TL_CHAIN_HALO=1: window 1 -- the frame in front of a block through chain 4's registers instead of overlapping boxes
    int chHalo = (cc.ns && cc.w0 == 1 && chSR == 1 && env_int("SOAP_TL_CHAIN_HALO", kChainHaloDefault) != 0) ? 1 : 0;
    if (chHalo) {
      if (ch_ok && ch_make_plan(dim, es, n_windows, chU, true, 0, true, cp)) chOV = 0;
      else chHalo = 0;
    }
    if (ch_ok && !chHalo && !ch_make_plan(dim, es, n_windows, chU, chSR != 0, chOV, false, cp)) {
      chOV = 0;
      ch_ok = ch_make_plan(dim, es, n_windows, chU, chSR != 0, chOV, false, cp);
    }
    if (ch_ok) {
#define CH_VAR(UU, SS)                                                                                              \
  if (chU == UU && chSR == SS)                                                                                      \
    rc = dtype == SOAP_F64 ? launch_chain_pattern<double, UU, SS>(X, mult, sc, n_frames, n_atoms, dim, cp, cc, chOV, chHalo, s) \
                           : launch_chain_pattern<float, UU, SS>(X, mult, sc, n_frames, n_atoms, dim, cp, cc, chOV, chHalo, s);
      rc = set_error(SOAP_EINVAL, "soap_timelag: no chain variant SOAP_TL_CHAIN_U=%d SOAP_TL_CHAIN_SR=%d", chU, chSR);
      CH_VAR(4, 4) CH_VAR(4, 1)
#undef CH_VAR
      if (rc) return rc;
      int64_t off[4] = {0, 0, 0, 0};
      for (int k = 0; k < n_windows; ++k) off[k] = t_offsets_host[cc.perm[k]];
      if (n_windows == 3) return launch_finalize<3>(sc, t, off, n_frames, n_atoms, 1, cc.w, kind, power, s);
      return launch_finalize<4>(sc, t, off, n_frames, n_atoms, 1, cc.w, kind, power, s);
    }
  }
  {
    const PairChoice pc = pair_choose(n_frames, n_windows, windows_host, kind, bulk_ok, es);
    PairPlan pp;
    if (pc.use && pr_make_plan(n_frames, dim, es, n_windows, pp)) {
#define PR_CASE(T, NW)                                                                                                      \
  rc = mult ? launch_timelag_pair<T, NW, true>(X, mult, sc, n_frames, n_atoms, dim, pp, pc, s)                              \
            : launch_timelag_pair<T, NW, false>(X, mult, sc, n_frames, n_atoms, dim, pp, pc, s)
      if (dtype == SOAP_F64) {
        if (n_windows == 3) PR_CASE(double, 3); else PR_CASE(double, 4);
      } else {
        if (n_windows == 3) PR_CASE(float, 3); else PR_CASE(float, 4);
      }
#undef PR_CASE
      if (rc) return rc;
      int64_t off[4] = {0, 0, 0, 0};
      for (int k = 0; k < n_windows; ++k) off[k] = t_offsets_host[pc.perm[k]];
      const int fsplit = tl_atom_major(n_atoms) ? 1 : pp.split;
      if (n_windows == 3) return launch_finalize<3>(sc, t, off, n_frames, n_atoms, fsplit, pc.w, kind, power, s);
      return launch_finalize<4>(sc, t, off, n_frames, n_atoms, fsplit, pc.w, kind, power, s);
    }
  }
  {
    const RbChoice ch = rb_choose(n_frames, n_atoms, dim, es, windows_host, n_windows, kind, bulk_ok != 0);
    RbPlan rp;
    if (ch.use && rb_make_plan(dim, es, n_windows, rp)) {
      rc = dtype == SOAP_F64 ? dispatch_timelag_rb<double>(X, mult, sc, n_frames, n_atoms, dim, rp, ch, s)
                             : dispatch_timelag_rb<float>(X, mult, sc, n_frames, n_atoms, dim, rp, ch, s);
      if (rc) return rc;
      int64_t off[4] = {0, 0, 0, 0};
      for (int k = 0; k < n_windows; ++k) off[k] = t_offsets_host[ch.perm[k]];
      switch (n_windows) {
        case 1: return launch_finalize<1>(sc, t, off, n_frames, n_atoms, 1, ch.w, kind, power, s);
        case 2: return launch_finalize<2>(sc, t, off, n_frames, n_atoms, 1, ch.w, kind, power, s);
        case 3: return launch_finalize<3>(sc, t, off, n_frames, n_atoms, 1, ch.w, kind, power, s);
        default: return launch_finalize<4>(sc, t, off, n_frames, n_atoms, 1, ch.w, kind, power, s);
      }
    }
  }
  TimelagPlan pl;
  if (!make_plan(n_frames, plan_dim, es, windows_host, n_windows, pl))
    return set_error(SOAP_EUNSUPPORTED, "soap_timelag: window %d needs more shared memory than one SM has", pl.max_lag);
  if (bulk_ok) bulk_ok |= env_int("SOAP_TL_DIAG", 0) & 30;
#define TL_CASE(T, NW)                                                                                          \
  case NW:                                                                                                      \
    rc = mult ? launch_timelag<T, NW, true>(X, mult, sc, n_frames, n_atoms, dim, pl, windows_host, kind, bulk_ok, mult_pitch, inv_norm, s)  \
              : launch_timelag<T, NW, false>(X, mult, sc, n_frames, n_atoms, dim, pl, windows_host, kind, bulk_ok, mult_pitch, inv_norm, s); \
    if (rc) return rc;                                                                                          \
    return launch_finalize<NW>(sc, t, t_offsets_host, n_frames, n_atoms, tl_atom_major(n_atoms) ? 1 : pl.split, windows_host,     \
                               kind == SOAP_KIND_EUCLID_NORMALIZED ? SOAP_KIND_EUCLID : kind, power, s);
  if (dtype == SOAP_F64) {
    switch (n_windows) { TL_CASE(double, 1) TL_CASE(double, 2) TL_CASE(double, 3) TL_CASE(double, 4) }
  } else {
    switch (n_windows) { TL_CASE(float, 1) TL_CASE(float, 2) TL_CASE(float, 3) TL_CASE(float, 4) }
  }
#undef TL_CASE
  return set_error(SOAP_EINVAL, "soap_timelag: n_windows %d", n_windows);
}

extern "C" int soap_diff_transposed(const double *t, double *dt, int64_t rows, int64_t n_atoms, void *stream) {
  SOAP_REQUIRE(rows >= 1 && n_atoms >= 1, "soap_diff_transposed: bad shape");
  if (rows == 1) return SOAP_OK;
  SOAP_REQUIRE(t && dt, "soap_diff_transposed: null buffer");
  dim3 grid((unsigned)((rows - 1 + 31) / 32), (unsigned)((n_atoms + 31) / 32));
  SOAP_REQUIRE(grid.y <= 65535, "soap_diff_transposed: too many atoms");
  diff_transposed_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(t, dt, (long long)rows,
                                                                              (long long)n_atoms);
  return check_launch("diff_transposed_kernel");
}
