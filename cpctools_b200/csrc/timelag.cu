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
// Partial sums (uu, uv_w) go to scratch[a][p][q][f]; a small finalize kernel sums the pieces in
// fixed order (deterministic), applies the distance epilogue and writes t_w[f-w][a].
#include "common.cuh"

#include <stdlib.h>

namespace soapb {

constexpr int kTlWarps = 12;                     // consumer warps (3 per SM sub-partition)
constexpr int kTlThreads = (kTlWarps + 2) * 32;  // + producer warp + reducer warp
constexpr int kBlk = 32;  // frames per block == lanes per warp

struct TimelagPlan {
  int unit_elems, n_units, split, upp, stride, R, ahead, Q, max_lag;
  size_t smem;
};

static size_t tl_smem_bytes(int R, int stride, int upp, int Q) {
  return (size_t)R * stride + (size_t)upp * 16 + (size_t)kTlWarps * Q * 32 * 8 + (size_t)(2 * (R / kBlk) + 2) * 8 + 64;
}

static int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

static bool make_plan(int D, int elem_size, const int *windows, int nw, TimelagPlan &pl) {
  pl.unit_elems = 16 / elem_size;
  pl.n_units = (D + pl.unit_elems - 1) / pl.unit_elems;
  pl.Q = 1 + nw;
  pl.max_lag = 0;
  for (int k = 0; k < nw; ++k) pl.max_lag = windows[k] > pl.max_lag ? windows[k] : pl.max_lag;
  const size_t budget = (size_t)max_smem_optin() - 1024;
  const int force_split = env_int("SOAP_TL_SPLIT", 0);    // tuning knobs (benchmarks only)
  const int force_ahead = env_int("SOAP_TL_AHEAD", 0);
  for (int split = force_split > 0 ? force_split : 1; split <= pl.n_units; ++split) {
    const int upp = (pl.n_units + split - 1) / split;
    const int real_split = (pl.n_units + upp - 1) / upp;  // no empty pieces
    const int stride = 16 * (upp | 1);
    for (int ahead = force_ahead > 0 ? force_ahead : 3; ahead >= 1; --ahead) {
      // ring = blocks still needed as lag partners + the block being consumed + `ahead` in flight
      const int R = kBlk * ((pl.max_lag + kBlk - 1) / kBlk + 1 + ahead);
      const size_t smem = tl_smem_bytes(R, stride, upp, pl.Q);
      if (smem <= budget) {
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

__device__ __forceinline__ void cp_async_elem(uint32_t dst, const void *src, int bytes) {
  if (bytes == 8)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t *bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// U units (16 bytes each) of the lane's frame against the NW lagged frames.  All loads are issued
// before the first use so that one warp keeps U*(NW+2) shared-memory requests in flight.
// fp64: exact products accumulated by DFMA, separate chains for the two halves of a unit.
template <int NW, int U, bool EUCLID>
__device__ __forceinline__ void accum_units_f64(uint32_t m_addr, uint32_t x_addr, const uint32_t (&y_addr)[NW],
                                                uint32_t off, uint32_t ustep, double (&acc)[2 * (NW + 1)]) {
  double2 m[U], x[U], y[NW][U];
#pragma unroll
  for (int j = 0; j < U; ++j) {
    m[j] = lds_d2(m_addr + off + j * ustep);
    x[j] = lds_d2(x_addr + off + j * ustep);
#pragma unroll
    for (int k = 0; k < NW; ++k) y[k][j] = lds_d2(y_addr[k] + off + j * ustep);
  }
#pragma unroll
  for (int j = 0; j < U; ++j) {
    const double xm0 = x[j].x * m[j].x, xm1 = x[j].y * m[j].y;
    acc[0] = fma(xm0, x[j].x, acc[0]);
    acc[1] = fma(xm1, x[j].y, acc[1]);
#pragma unroll
    for (int k = 0; k < NW; ++k) {
      if (EUCLID) {
        const double d0 = x[j].x - y[k][j].x, d1 = x[j].y - y[k][j].y;
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
template <int NW, int U, bool EUCLID>
__device__ __forceinline__ void accum_units_f32(uint32_t m_addr, uint32_t x_addr, const uint32_t (&y_addr)[NW],
                                                uint32_t off, uint32_t ustep, double (&acc)[2 * (NW + 1)]) {
  float4 m[U], x[U], y[NW][U];
#pragma unroll
  for (int j = 0; j < U; ++j) {
    m[j] = lds_f4(m_addr + off + j * ustep);
    x[j] = lds_f4(x_addr + off + j * ustep);
#pragma unroll
    for (int k = 0; k < NW; ++k) y[k][j] = lds_f4(y_addr[k] + off + j * ustep);
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
        const float d0 = x[j].x - y[k][j].x, d1 = x[j].y - y[k][j].y, d2 = x[j].z - y[k][j].z, d3 = x[j].w - y[k][j].w;
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

template <typename T, int NW, int U, bool EUCLID>
__device__ __forceinline__ void accum_units(uint32_t m_addr, uint32_t x_addr, const uint32_t (&y_addr)[NW],
                                            uint32_t off, uint32_t ustep, double (&acc)[2 * (NW + 1)]) {
  if constexpr (sizeof(T) == 8)
    accum_units_f64<NW, U, EUCLID>(m_addr, x_addr, y_addr, off, ustep, acc);
  else
    accum_units_f32<NW, U, EUCLID>(m_addr, x_addr, y_addr, off, ustep, acc);
}

// Warp roles: warps 0..11 consume (lane = frame), warp 12 produces (TMA / cp.async), warp 13 reduces
// the consumers' partial sums of a block and writes them to scratch.  All hand-offs are
// mbarriers; there is no __syncthreads after the prologue, so no warp ever convoys behind another.
//   full[s]      producer -> consumers   block in ring slot-group s has landed
//   empty[s]     consumers -> producer   block s consumed as "current" frames (one arrival per warp)
//   red_full     consumers -> reducer    partial sums of a block are in red (one arrival per warp)
//   red_empty    reducer -> consumers    red may be overwritten
template <typename T, int NW, bool EUCLID>
__global__ void __launch_bounds__(kTlThreads, 1)
    timelag_kernel(const T *__restrict__ X, const T *__restrict__ mult, double *__restrict__ scratch,
                   long long F, long long A, int D, int split, int upp, int stride, int R, int lag_blocks,
                   int4 win4, int bulk_ok) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int UE = 16 / (int)sizeof(T);
  constexpr int Q = NW + 1;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int win[4] = {win4.x, win4.y, win4.z, win4.w};
  const int nslots = R / kBlk;

  unsigned char *ring = smem_raw;
  unsigned char *msm = ring + (size_t)R * stride;
  double *red = reinterpret_cast<double *>(msm + (size_t)upp * 16);      // [kTlWarps][Q][32]
  uint64_t *full = reinterpret_cast<uint64_t *>(red + kTlWarps * Q * 32);  // [nslots]
  uint64_t *empty = full + nslots;                                        // [nslots]
  uint64_t *red_full = empty + nslots;                                    // [1]
  uint64_t *red_empty = red_full + 1;                                     // [1]

  const long long a = blockIdx.x / split;
  const int p = (int)(blockIdx.x - a * split);
  const int e0 = p * upp * UE;
  const int pe = min(D - e0, upp * UE);  // elements of this piece (> 0)
  const int punits = (pe + UE - 1) / UE;
  const uint32_t pbytes = (uint32_t)pe * sizeof(T);
  const int nblk = (int)((F + kBlk - 1) / kBlk);

  // prologue (all warps): multiplicities of this piece, zero padded; ring zeroed when rows are not
  // 16-byte multiples (the tail unit is then partial and its padding must read as zero)
  for (int i = tid; i < upp * UE; i += kTlThreads)
    reinterpret_cast<T *>(msm)[i] = (i < pe) ? (mult ? mult[e0 + i] : (T)1) : (T)0;
  if (!bulk_ok) {
    const int n16 = (int)(((size_t)R * stride) / 16);
    for (int i = tid; i < n16; i += kTlThreads) reinterpret_cast<int4 *>(ring)[i] = make_int4(0, 0, 0, 0);
  }
  if (tid == 0) {
    for (int s = 0; s < nslots; ++s) {
      mbar_init(&full[s], bulk_ok ? 1 : 32);
      mbar_init(&empty[s], kTlWarps);
    }
    mbar_init(red_full, kTlWarps);
    mbar_init(red_empty, 1);
    fence_mbar_init();
  }
  __syncthreads();

  if (warp == kTlWarps) {
    // ===== producer ==========================================================================
    const size_t fstride = (size_t)A * D;  // elements between consecutive frames
    const T *xa = X + (size_t)a * D + e0;
    const uint32_t ring_addr = smem_u32(ring);
    for (int c = 0; c < nblk; ++c) {
      // loading block c overwrites the frames of block c - nslots, last needed (as lag partners)
      // by block j = c - nslots + lag_blocks: wait until the consumers are done with block j
      const int j = c - nslots + lag_blocks;
      if (j >= 0) mbar_wait(&empty[j % nslots], (uint32_t)((j / nslots) & 1));
      const long long f0 = (long long)c * kBlk;
      const int nf = (int)min((long long)kBlk, F - f0);
      const int s = c % nslots;
      if (bulk_ok) {
        if (lane == 0) mbar_expect_tx(&full[s], (uint32_t)nf * pbytes);
        __syncwarp();
        if (lane < nf)
          bulk_g2s(ring + (size_t)(s * kBlk + lane) * stride, xa + (size_t)(f0 + lane) * fstride, pbytes, &full[s]);
      } else {
        for (int fr = 0; fr < nf; ++fr) {
          const T *src = xa + (size_t)(f0 + fr) * fstride;
          const uint32_t dst = ring_addr + (uint32_t)(s * kBlk + fr) * stride;
          for (int e = lane; e < pe; e += 32) cp_async_elem(dst + e * (uint32_t)sizeof(T), src + e, (int)sizeof(T));
        }
        cp_async_arrive_noinc(&full[s]);  // every lane: arrives once its own copies have landed
      }
    }
  } else if (warp == kTlWarps + 1) {
    // ===== reducer ===========================================================================
    double *srow = scratch + ((size_t)a * split + p) * Q * (size_t)F;
    for (int b = 0; b < nblk; ++b) {
      mbar_wait(red_full, (uint32_t)(b & 1));
      const double *r = red;
      const long long fl = (long long)b * kBlk + lane;
#pragma unroll
      for (int q = 0; q < Q; ++q) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kTlWarps; ++w) s += r[(w * Q + q) * 32 + lane];
        if (fl < F) srow[(size_t)q * F + fl] = s;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(red_empty);
    }
  } else {
    // ===== consumers: lane <-> frame ============================================================
    const uint32_t ring_addr = smem_u32(ring), m_addr = smem_u32(msm);
    const uint32_t ustep = 16u * kTlWarps;  // this warp's units are u = warp, warp + kTlWarps, ...
    const uint32_t off0 = 16u * warp;
    const int my_units = punits > warp ? (punits - warp + kTlWarps - 1) / kTlWarps : 0;
    int wmod[NW];
#pragma unroll
    for (int k = 0; k < NW; ++k) wmod[k] = win[k] % R;
    for (int b = 0; b < nblk; ++b) {
      const int s = b % nslots;
      const int slot = s * kBlk + lane;  // ring slot of this lane's frame (R is a multiple of 32)
      const uint32_t x_addr = ring_addr + (uint32_t)slot * stride;
      uint32_t y_addr[NW];
#pragma unroll
      for (int k = 0; k < NW; ++k) {
        int ys = slot - wmod[k];
        if (ys < 0) ys += R;
        y_addr[k] = ring_addr + (uint32_t)ys * stride;  // frames f < w read stale data: result unused
      }
      double acc[2 * Q];
#pragma unroll
      for (int q = 0; q < 2 * Q; ++q) acc[q] = 0.0;
      mbar_wait(&full[s], (uint32_t)((b / nslots) & 1));
      int n = 0;
      for (; n + 4 <= my_units; n += 4)
        accum_units<T, NW, 4, EUCLID>(m_addr, x_addr, y_addr, off0 + n * ustep, ustep, acc);
      if (n + 2 <= my_units) {
        accum_units<T, NW, 2, EUCLID>(m_addr, x_addr, y_addr, off0 + n * ustep, ustep, acc);
        n += 2;
      }
      if (n < my_units) accum_units<T, NW, 1, EUCLID>(m_addr, x_addr, y_addr, off0 + n * ustep, ustep, acc);
      // hand the partial sums to the reducer (it is ~10x faster than a block: this wait is free)
      if (b >= 1) mbar_wait(red_empty, (uint32_t)((b - 1) & 1));
      double *r = red + (size_t)warp * Q * 32;
#pragma unroll
      for (int q = 0; q < Q; ++q) r[q * 32 + lane] = acc[2 * q] + acc[2 * q + 1];
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(red_full);
        mbar_arrive(&empty[s]);
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
      double uu = 0.0;
      for (int p = 0; p < split; ++p) uu += base[((size_t)p * Q) * F + f];
#pragma unroll
      for (int k = 0; k < NW; ++k) {
        if (f >= win[k]) {
          double uv = 0.0, vv = 0.0;
          for (int p = 0; p < split; ++p) {
            uv += base[((size_t)p * Q + 1 + k) * F + f];
            vv += base[((size_t)p * Q) * F + (f - win[k])];
          }
          tile[k][lx][al] = soap_epilogue(uv, uu, vv, kind, power);
        }
      }
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

template <typename T, int NW>
static int launch_timelag(const void *X, const void *mult, double *scratch, int64_t F, int64_t A, int D,
                          const TimelagPlan &pl, const int *w, int kind, int bulk_ok, cudaStream_t s) {
  const int4 win4 = make_int4(w[0], NW > 1 ? w[1] : 0, NW > 2 ? w[2] : 0, NW > 3 ? w[3] : 0);
  auto kern = (kind == SOAP_KIND_EUCLID) ? timelag_kernel<T, NW, true> : timelag_kernel<T, NW, false>;
  SOAP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  const long long nblocks = (long long)A * pl.split;
  SOAP_REQUIRE(nblocks < 2147483647LL, "soap_timelag: atoms x pieces = %lld exceeds the grid limit", nblocks);
  ProfScope prof(SOAP_PROF_TIMELAG, s);
  kern<<<(unsigned)nblocks, kTlThreads, pl.smem, s>>>(static_cast<const T *>(X), static_cast<const T *>(mult),
                                                      scratch, (long long)F, (long long)A, D, pl.split, pl.upp,
                                                      pl.stride, pl.R, (pl.max_lag + kBlk - 1) / kBlk, win4, bulk_ok);
  return check_launch("timelag_kernel");
}

template <int NW>
static int launch_finalize(const double *scratch, double *t, const int64_t *off, int64_t F, int64_t A,
                           const TimelagPlan &pl, const int *w, int kind, int power, cudaStream_t s) {
  const int4 win4 = make_int4(w[0], NW > 1 ? w[1] : 0, NW > 2 ? w[2] : 0, NW > 3 ? w[3] : 0);
  dim3 grid((unsigned)((F + 31) / 32), (unsigned)((A + 31) / 32));
  SOAP_REQUIRE(grid.y <= 65535, "soap_timelag: more than 2M atoms per call are not supported");
  timelag_finalize_kernel<NW><<<grid, 256, 0, s>>>(scratch, t, off[0], NW > 1 ? off[1] : 0, NW > 2 ? off[2] : 0,
                                                  NW > 3 ? off[3] : 0, (long long)F, (long long)A, pl.split,
                                                  win4, kind, power);
  return check_launch("timelag_finalize_kernel");
}

}  // namespace soapb

using namespace soapb;

static int check_windows(int64_t F, const int *w, int nw, const char *who) {
  SOAP_REQUIRE(w != nullptr && nw >= 1 && nw <= SOAP_MAX_WINDOWS, "%s: 1..%d windows per call", who, SOAP_MAX_WINDOWS);
  for (int k = 0; k < nw; ++k)
    SOAP_REQUIRE(w[k] >= 1 && w[k] < F, "%s: window %d must be in [1, n_frames)", who, w[k]);
  return SOAP_OK;
}

extern "C" size_t soap_timelag_scratch_bytes(int64_t n_frames, int64_t n_atoms, int dim, int dtype,
                                             const int *windows_host, int n_windows) {
  if (n_frames <= 0 || n_atoms <= 0 || dim <= 0 || !windows_host || n_windows < 1 || n_windows > SOAP_MAX_WINDOWS)
    return 0;
  TimelagPlan pl;
  if (!make_plan(dim, dtype == SOAP_F64 ? 8 : 4, windows_host, n_windows, pl)) return 0;
  return (size_t)n_atoms * pl.split * pl.Q * (size_t)n_frames * sizeof(double);
}

extern "C" int soap_timelag(const void *X, const void *mult, double *t, const int64_t *t_offsets_host,
                            int64_t n_frames, int64_t n_atoms, int dim, const int *windows_host,
                            int n_windows, int kind, int power, int dtype, void *scratch, void *stream) {
  SOAP_REQUIRE(n_frames > 0 && n_atoms > 0 && dim > 0, "soap_timelag: bad shape");
  SOAP_REQUIRE(X && t && t_offsets_host && scratch, "soap_timelag: null buffer");
  SOAP_REQUIRE(dtype == SOAP_F32 || dtype == SOAP_F64, "soap_timelag: dtype %d", dtype);
  SOAP_REQUIRE(kind >= SOAP_KIND_SIMPLE && kind <= SOAP_KIND_NORMALIZED_FUSED, "soap_timelag: kind %d", kind);
  int rc = check_windows(n_frames, windows_host, n_windows, "soap_timelag");
  if (rc) return rc;
  const int es = dtype == SOAP_F64 ? 8 : 4;
  TimelagPlan pl;
  if (!make_plan(dim, es, windows_host, n_windows, pl))
    return set_error(SOAP_EUNSUPPORTED, "soap_timelag: window %d needs more shared memory than one SM has", pl.max_lag);
  const int bulk_ok = ((reinterpret_cast<uintptr_t>(X) & 15) == 0) && (((size_t)dim * es) % 16 == 0);
  auto s = static_cast<cudaStream_t>(stream);
  auto sc = static_cast<double *>(scratch);
#define TL_CASE(T, NW)                                                                                          \
  case NW:                                                                                                      \
    rc = launch_timelag<T, NW>(X, mult, sc, n_frames, n_atoms, dim, pl, windows_host, kind, bulk_ok, s);        \
    if (rc) return rc;                                                                                          \
    return launch_finalize<NW>(sc, t, t_offsets_host, n_frames, n_atoms, pl, windows_host, kind, power, s);
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
