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

constexpr int kTlThreads = 256;
constexpr int kTlWarps = kTlThreads / 32;
constexpr int kBlk = 32;  // frames per block == lanes per warp

struct TimelagPlan {
  int unit_elems, n_units, split, upp, stride, R, ahead, Q, max_lag;
  size_t smem;
};

static size_t tl_smem_bytes(int R, int stride, int upp, int Q) {
  return (size_t)R * stride + (size_t)upp * 16 + (size_t)kTlWarps * Q * 32 * 8 + (size_t)(R / kBlk) * 8 + 64;
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
      const int R = ((pl.max_lag + kBlk * (ahead + 1) + kBlk - 1) / kBlk) * kBlk;
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

template <typename T>
struct Vec16;
template <>
struct Vec16<double> { using type = double2; };
template <>
struct Vec16<float> { using type = float4; };

// one 16-byte unit of frame x against NW lagged frames, accumulated in double
template <int NW, bool EUCLID>
__device__ __forceinline__ void accum_unit(const double2 m, const double2 x, const double2 (&y)[NW],
                                           double (&acc)[NW + 1]) {
  const double xm0 = x.x * m.x, xm1 = x.y * m.y;
  acc[0] = fma(xm0, x.x, acc[0]);
  acc[0] = fma(xm1, x.y, acc[0]);
#pragma unroll
  for (int k = 0; k < NW; ++k) {
    if (EUCLID) {
      const double d0 = x.x - y[k].x, d1 = x.y - y[k].y;
      acc[k + 1] = fma(m.x * d0, d0, acc[k + 1]);
      acc[k + 1] = fma(m.y * d1, d1, acc[k + 1]);
    } else {
      acc[k + 1] = fma(xm0, y[k].x, acc[k + 1]);
      acc[k + 1] = fma(xm1, y[k].y, acc[k + 1]);
    }
  }
}

// fp32 data: products and an 8-term chain in fp32, chain sums promoted to double (keeps the
// conversion count at 1/8 per product while bounding the fp32 rounding to one short chain)
template <int NW, bool EUCLID>
__device__ __forceinline__ void accum_unit2(const float4 ma, const float4 xa, const float4 (&ya)[NW],
                                            const float4 mb, const float4 xb, const float4 (&yb)[NW],
                                            double (&acc)[NW + 1]) {
  const float a0 = xa.x * ma.x, a1 = xa.y * ma.y, a2 = xa.z * ma.z, a3 = xa.w * ma.w;
  const float b0 = xb.x * mb.x, b1 = xb.y * mb.y, b2 = xb.z * mb.z, b3 = xb.w * mb.w;
  float c = a0 * xa.x;
  c = fmaf(a1, xa.y, c); c = fmaf(a2, xa.z, c); c = fmaf(a3, xa.w, c);
  c = fmaf(b0, xb.x, c); c = fmaf(b1, xb.y, c); c = fmaf(b2, xb.z, c); c = fmaf(b3, xb.w, c);
  acc[0] += (double)c;
#pragma unroll
  for (int k = 0; k < NW; ++k) {
    float s;
    if (EUCLID) {
      float d;
      d = xa.x - ya[k].x; s = ma.x * d * d;
      d = xa.y - ya[k].y; s = fmaf(ma.y * d, d, s);
      d = xa.z - ya[k].z; s = fmaf(ma.z * d, d, s);
      d = xa.w - ya[k].w; s = fmaf(ma.w * d, d, s);
      d = xb.x - yb[k].x; s = fmaf(mb.x * d, d, s);
      d = xb.y - yb[k].y; s = fmaf(mb.y * d, d, s);
      d = xb.z - yb[k].z; s = fmaf(mb.z * d, d, s);
      d = xb.w - yb[k].w; s = fmaf(mb.w * d, d, s);
    } else {
      s = a0 * ya[k].x;
      s = fmaf(a1, ya[k].y, s); s = fmaf(a2, ya[k].z, s); s = fmaf(a3, ya[k].w, s);
      s = fmaf(b0, yb[k].x, s); s = fmaf(b1, yb[k].y, s); s = fmaf(b2, yb[k].z, s); s = fmaf(b3, yb[k].w, s);
    }
    acc[k + 1] += (double)s;
  }
}

template <typename T, int NW, bool EUCLID>
__global__ void __launch_bounds__(kTlThreads, 1)
    timelag_kernel(const T *__restrict__ X, const T *__restrict__ mult, double *__restrict__ scratch,
                   long long F, long long A, int D, int split, int upp, int stride, int R, int ahead,
                   int4 win4, int bulk_ok) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int UE = 16 / (int)sizeof(T);
  constexpr int Q = NW + 1;
  using V = typename Vec16<T>::type;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int win[4] = {win4.x, win4.y, win4.z, win4.w};

  unsigned char *ring = smem_raw;
  unsigned char *msm = ring + (size_t)R * stride;
  double *red = reinterpret_cast<double *>(msm + (size_t)upp * 16);
  uint64_t *bars = reinterpret_cast<uint64_t *>(red + kTlWarps * Q * 32);
  const int nslots = R / kBlk;

  const long long a = blockIdx.x / split;
  const int p = (int)(blockIdx.x - a * split);
  const int e0 = p * upp * UE;
  const int pe = min(D - e0, upp * UE);          // elements of this piece (> 0)
  const int punits = (pe + UE - 1) / UE;
  const uint32_t pbytes = (uint32_t)pe * sizeof(T);

  // multiplicities of this piece (zero padded), ring zeroed when the tail unit is partial
  for (int i = tid; i < upp * UE; i += kTlThreads)
    reinterpret_cast<T *>(msm)[i] = (i < pe) ? (mult ? mult[e0 + i] : (T)1) : (T)0;
  if (!bulk_ok) {
    const int n16 = (int)(((size_t)R * stride) / 16);
    for (int i = tid; i < n16; i += kTlThreads) reinterpret_cast<int4 *>(ring)[i] = make_int4(0, 0, 0, 0);
  }
  if (tid == 0) {
    for (int s = 0; s < nslots; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  __syncthreads();

  const size_t fstride = (size_t)A * D;  // elements between consecutive frames
  const T *xa = X + (size_t)a * D + e0;
  const int nblk = (int)((F + kBlk - 1) / kBlk);
  double *srow = scratch + ((size_t)a * split + p) * Q * (size_t)F;
  int issued = 0;

  for (int b = 0; b < nblk; ++b) {
    const int want = min(nblk - 1, b + ahead);
    if (bulk_ok) {
      if (warp == 0) {
        // all warps passed the barrier that ended block b-1: the slots of blocks <= want are free
        while (issued <= want) {
          const long long f0 = (long long)issued * kBlk;
          const int nf = (int)min((long long)kBlk, F - f0);
          uint64_t *bar = &bars[issued % nslots];
          if (lane == 0) mbar_expect_tx(bar, (uint32_t)nf * pbytes);
          __syncwarp();
          if (lane < nf) {
            const long long f = f0 + lane;
            bulk_g2s(ring + (size_t)(f % R) * stride, xa + (size_t)f * fstride, pbytes, bar);
          }
          ++issued;
        }
      }
      mbar_wait(&bars[b % nslots], (uint32_t)((b / nslots) & 1));
    } else {
      while (issued <= want) {  // unaligned rows: cooperative element-wise staging
        const long long f0 = (long long)issued * kBlk;
        const int nf = (int)min((long long)kBlk, F - f0);
        for (int i = tid; i < nf * pe; i += kTlThreads) {
          const int fr = i / pe, e = i - fr * pe;
          const long long f = f0 + fr;
          reinterpret_cast<T *>(ring + (size_t)(f % R) * stride)[e] = xa[(size_t)f * fstride + e];
        }
        ++issued;
      }
      __syncthreads();
    }

    // ---- compute: lane <-> frame fl ------------------------------------------------------
    const long long fl = (long long)b * kBlk + lane;
    const unsigned char *xb = ring + (size_t)(fl % R) * stride;
    const unsigned char *yb[NW];
#pragma unroll
    for (int k = 0; k < NW; ++k) {
      long long fp = fl - win[k];
      if (fp < 0) fp = fl;  // no partner yet: result unused, keep the address in range
      yb[k] = ring + (size_t)(fp % R) * stride;
    }
    double acc[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) acc[q] = 0.0;

    if constexpr (sizeof(T) == 8) {
#pragma unroll 2
      for (int u = warp; u < punits; u += kTlWarps) {
        const double2 m = *reinterpret_cast<const double2 *>(msm + 16 * u);
        const double2 x = *reinterpret_cast<const double2 *>(xb + 16 * u);
        double2 y[NW];
#pragma unroll
        for (int k = 0; k < NW; ++k) y[k] = *reinterpret_cast<const double2 *>(yb[k] + 16 * u);
        accum_unit<NW, EUCLID>(m, x, y, acc);
      }
    } else {
      for (int u = warp; u < punits; u += 2 * kTlWarps) {
        const int u2 = u + kTlWarps;
        const bool two = u2 < punits;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 ma = *reinterpret_cast<const float4 *>(msm + 16 * u);
        const float4 xa4 = *reinterpret_cast<const float4 *>(xb + 16 * u);
        const float4 mb = two ? *reinterpret_cast<const float4 *>(msm + 16 * u2) : z;
        const float4 xb4 = two ? *reinterpret_cast<const float4 *>(xb + 16 * u2) : z;
        float4 ya[NW], yb4[NW];
#pragma unroll
        for (int k = 0; k < NW; ++k) {
          ya[k] = *reinterpret_cast<const float4 *>(yb[k] + 16 * u);
          yb4[k] = two ? *reinterpret_cast<const float4 *>(yb[k] + 16 * u2) : z;
        }
        accum_unit2<NW, EUCLID>(ma, xa4, ya, mb, xb4, yb4, acc);
      }
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) red[(warp * Q + q) * 32 + lane] = acc[q];
    __syncthreads();
    if (tid < Q * 32) {
      const int q = tid >> 5;
      double s = 0.0;
#pragma unroll
      for (int w = 0; w < kTlWarps; ++w) s += red[(w * Q + q) * 32 + lane];
      if (fl < F) srow[(size_t)q * F + fl] = s;
    }
    __syncthreads();  // ends block b: ring slots older than max_lag and `red` may be reused
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
                                                      pl.stride, pl.R, pl.ahead, win4, bulk_ok);
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
