// expand.cu -- K1/K2: gather-expand of compressed SOAP power spectra, optionally fused with
// the L2 normalisation of the expanded vector.
//
// Reference semantics: fillSOAPVectorFromdscribe (src/SOAPify/utils.py:271-324: x[..., idx]) and
// normalizeArray (utils.py:140-153: x / norm, zero-norm rows left at zero).
//
// Memory-bound (reads Dc, writes Df elements per vector).  One CTA owns a tile of TV
// consecutive vectors -- contiguous in memory -- which the TMA engine drops into shared
// memory with ONE 1-D bulk copy (cp.async.bulk, double buffered, mbarrier completion); the
// gather then runs out of shared memory and leaves as fully coalesced 128-bit streaming stores.
// The int32 gather table lives in shared memory and is hoisted into registers per column chunk.
#include "common.cuh"

#include <stdlib.h>

#include <type_traits>

namespace soapb {

template <typename T>
struct VecOf;  // 16-byte vector of T
template <>
struct VecOf<double> { static constexpr int N = 2; };
template <>
struct VecOf<float> { static constexpr int N = 4; };
template <>
struct VecOf<uint64_t> { static constexpr int N = 2; };
template <>
struct VecOf<uint32_t> { static constexpr int N = 4; };

constexpr int kExpandThreads = 256;
constexpr int kMaxSplit = 4;  // warps that may share one vector in the norm phase

// x / norm in the data's own precision, as numpy computes it.  The quotient is formed from the
// row's reciprocal with one FMA correction step (q = v*r; q += fma(-q, n, v) * r), which is the
// correctly rounded IEEE quotient for normal operands at 3 instructions per element instead of a
// full division sequence; rows whose norm is not a normal number take the plain division (rcp == 0).
template <typename T>
__device__ __forceinline__ T scale_elem(T v, double nrm, double rcp) {
  return v;
}
template <>
__device__ __forceinline__ double scale_elem<double>(double v, double nrm, double rcp) {
  if (rcp == 0.0) return v / nrm;
  const double q = v * rcp;
  return fma(fma(-q, nrm, v), rcp, q);
}
template <>
__device__ __forceinline__ float scale_elem<float>(float v, double nrm, double rcp) {
  const float n = (float)nrm, r = (float)rcp;
  if (r == 0.f) return v / n;
  const float q = v * r;
  return fmaf(fmaf(-q, n, v), r, q);
}

template <typename T>
__device__ __forceinline__ unsigned long long pun64(T v) {
  static_assert(sizeof(T) == 8, "");
  return *reinterpret_cast<unsigned long long *>(&v);
}
template <typename T>
__device__ __forceinline__ unsigned int pun32(T v) {
  static_assert(sizeof(T) == 4, "");
  return *reinterpret_cast<unsigned int *>(&v);
}

template <typename T>
__device__ __forceinline__ float to_float(T v) {
  return 0.f;
}
template <>
__device__ __forceinline__ float to_float<float>(float v) {
  return v;
}

template <typename T>
__device__ __forceinline__ double sq_as_double(T v) {
  return 0.0;
}
template <>
__device__ __forceinline__ double sq_as_double<double>(double v) {
  return v * v;
}
template <>
__device__ __forceinline__ double sq_as_double<float>(float v) {
  return (double)v * (double)v;
}

template <typename T>
__device__ __forceinline__ double weighted_sq(T v, T w) {
  return 0.0;
}
template <>
__device__ __forceinline__ double weighted_sq<double>(double v, double w) {
  return v * w * v;
}
template <>
__device__ __forceinline__ double weighted_sq<float>(float v, float w) {
  return (double)v * (double)w * (double)v;
}

// T: element type; NORM: fuse normalisation; HAS_IDX: gather table present (else identity)
template <typename T, bool NORM, bool HAS_IDX>
__global__ void __launch_bounds__(kExpandThreads, 2)
    expand_kernel(const T *__restrict__ in, T *__restrict__ out, const int32_t *__restrict__ idx,
                  long long nvec, int in_stride, int d_full, int tv, int bulk_ok, int vec_ok) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  constexpr int V = VecOf<T>::N;
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kExpandThreads / 32;

  const size_t stage_elems = (size_t)tv * in_stride;
  const size_t stage_bytes = ((stage_elems * sizeof(T)) + 127) & ~(size_t)127;
  auto stage = [&](int i) { return reinterpret_cast<T *>(smem_raw + (size_t)i * stage_bytes); };
  int32_t *tab = reinterpret_cast<int32_t *>(smem_raw + 2 * stage_bytes);
  const int d_pad = (d_full + 3) & ~3;
  // fp32 rows take their squared norm from the COMPRESSED row with slot weights (4x fewer, contiguous
  // 128-bit loads); fp64 rows are faster through the gather (measured: 0.996 vs 0.93 of the HBM peak)
  constexpr bool WEIGHTED = NORM && HAS_IDX && sizeof(T) == 4;
  const int in_pad = (in_stride + 3) & ~3;
  T *wm = reinterpret_cast<T *>(tab + (HAS_IDX ? d_pad : 0));  // [in_pad] slot weights, 16-byte aligned
  double *nrm = reinterpret_cast<double *>(wm + (WEIGHTED ? in_pad : 0));
  double *rcp = nrm + ((tv + 1) & ~1);    // reciprocal norms (0: take the division path)
  double *parts = rcp + ((tv + 1) & ~1);  // [tv][kMaxSplit] partial square sums
  uint64_t *bar = reinterpret_cast<uint64_t *>(parts + (size_t)tv * kMaxSplit);

  const long long ntiles = (nvec + tv - 1) / tv;
  if (HAS_IDX) {
    for (int j = tid; j < d_pad; j += kExpandThreads) tab[j] = (j < d_full) ? idx[j] : 0;
  }
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  if constexpr (WEIGHTED) {
    // weights: occurrences of every compressed slot in the table, counted in place as integers
    for (int j = tid; j < in_pad; j += kExpandThreads) wm[j] = T(0);
    __syncthreads();
    for (int j = tid; j < d_full; j += kExpandThreads) atomicAdd(reinterpret_cast<int *>(wm + idx[j]), 1);
    __syncthreads();
    for (int j = tid; j < in_pad; j += kExpandThreads) wm[j] = (T)(*reinterpret_cast<const int *>(wm + j));
  }
  __syncthreads();

  auto tile_rows = [&](long long t) -> int {
    long long left = nvec - t * tv;
    return (int)(left < tv ? left : tv);
  };
  auto issue = [&](long long t, int s) {  // thread 0 only, bulk path
    const uint32_t bytes = (uint32_t)((size_t)tile_rows(t) * in_stride * sizeof(T));
    mbar_expect_tx(&bar[s], bytes);
    bulk_g2s(stage(s), in + (size_t)t * tv * in_stride, bytes, &bar[s]);
  };

  // bulk_ok bit 1: tiles are staged by bulk copies -- rows need not be 16-byte multiples for that, only TILES (the
  // host picks tv so that a full tile is; a ragged last tile whose byte count is not falls back to the plain copy);
  // bit 0: every row is a whole number of 16-byte chunks (vector reads inside the stage)
  const bool stage_bulk = (bulk_ok & 2) != 0;
  bulk_ok &= 1;
  auto tile_bulk = [&](long long tt) { return stage_bulk && (((size_t)tile_rows(tt) * in_stride * sizeof(T)) & 15) == 0; };
  long long t = blockIdx.x;
  int s = 0;
  uint32_t phase_bits = 0;  // bit s = parity of the next completion of bar[s]
  if (tid == 0 && t < ntiles && tile_bulk(t)) issue(t, 0);

  for (; t < ntiles; t += gridDim.x, s ^= 1) {
    const int rows = tile_rows(t);
    const long long tn = t + gridDim.x;
    // stage s^1 was last read in the previous iteration (trailing __syncthreads below)
    if (tid == 0 && tn < ntiles && tile_bulk(tn)) issue(tn, s ^ 1);
    if (tile_bulk(t)) {
      mbar_wait(&bar[s], (phase_bits >> s) & 1u);
      phase_bits ^= 1u << s;
    } else {
      const T *src = in + (size_t)t * tv * in_stride;
      const int n = rows * in_stride;
      for (int i = tid; i < n; i += kExpandThreads) stage(s)[i] = src[i];
      __syncthreads();
    }
    const T *sm = stage(s);

    if (NORM) {
      // phase 1: squared norm of the EXPANDED vector: a gather through the table, or (WEIGHTED) the
      // compressed row as sum_j w_j x_j^2 with w_j = how often slot j occurs in the table.  Each
      // vector is summed by `32/L * split` lane groups: half warps
      // when the tile holds more vectors than warps (short rows), several warps per vector when it
      // holds fewer (long rows).
      const int L = rows > kWarps ? 16 : 32;                       // lanes per group
      const int slots = kWarps * (32 / L);                         // groups available per round
      const int split = rows < slots ? (slots / rows < kMaxSplit ? slots / rows : kMaxSplit) : 1;
      const int grp = warp * (32 / L) + lane / L, sl = lane % L;
      for (int v0 = 0; v0 < rows; v0 += slots / split) {
        const int v = v0 + grp / split, part = grp % split;
        double acc = 0.0;
        if (v < rows && grp / split < slots / split) {
          if constexpr (WEIGHTED) {
            const T *row = sm + (size_t)v * in_stride;
            const int step = L * split;
            if (bulk_ok) {  // rows are whole 16-byte chunks and 16-byte aligned in the stage
              const int nch = in_stride / V;
              int c = sl + part * L;
              if constexpr (sizeof(T) == 4) {
                float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;  // fp32 partials over a slice of the row
                for (; c + step < nch; c += 2 * step) {
                  const float4 x0 = *reinterpret_cast<const float4 *>(row + 4 * c), w0 = *reinterpret_cast<const float4 *>(wm + 4 * c);
                  const float4 x1 = *reinterpret_cast<const float4 *>(row + 4 * (c + step)),
                               w1 = *reinterpret_cast<const float4 *>(wm + 4 * (c + step));
                  p0 = fmaf(x0.x * w0.x, x0.x, p0); p1 = fmaf(x0.y * w0.y, x0.y, p1);
                  p2 = fmaf(x0.z * w0.z, x0.z, p2); p3 = fmaf(x0.w * w0.w, x0.w, p3);
                  p0 = fmaf(x1.x * w1.x, x1.x, p0); p1 = fmaf(x1.y * w1.y, x1.y, p1);
                  p2 = fmaf(x1.z * w1.z, x1.z, p2); p3 = fmaf(x1.w * w1.w, x1.w, p3);
                }
                for (; c < nch; c += step) {
                  const float4 x0 = *reinterpret_cast<const float4 *>(row + 4 * c), w0 = *reinterpret_cast<const float4 *>(wm + 4 * c);
                  p0 = fmaf(x0.x * w0.x, x0.x, p0); p1 = fmaf(x0.y * w0.y, x0.y, p1);
                  p2 = fmaf(x0.z * w0.z, x0.z, p2); p3 = fmaf(x0.w * w0.w, x0.w, p3);
                }
                acc = ((double)p0 + (double)p1) + ((double)p2 + (double)p3);
              }
            } else {
              double a0 = 0.0, a1 = 0.0;
              int j = sl + part * L;
              for (; j + step < in_stride; j += 2 * step) {
                a0 += weighted_sq<T>(row[j], wm[j]);
                a1 += weighted_sq<T>(row[j + step], wm[j + step]);
              }
              for (; j < in_stride; j += step) a0 += weighted_sq<T>(row[j], wm[j]);
              acc = a0 + a1;
            }
          } else {
            const T *row = sm + (size_t)v * in_stride;
            const int step = L * split;
            int j = sl + part * L;
            if constexpr (sizeof(T) == 4) {
              float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;  // fp32 partials over a slice of the row
              for (; j + 3 * step < d_full; j += 4 * step) {
                const float e0 = to_float(row[HAS_IDX ? tab[j] : j]), e1 = to_float(row[HAS_IDX ? tab[j + step] : j + step]);
                const float e2 = to_float(row[HAS_IDX ? tab[j + 2 * step] : j + 2 * step]);
                const float e3 = to_float(row[HAS_IDX ? tab[j + 3 * step] : j + 3 * step]);
                p0 = fmaf(e0, e0, p0); p1 = fmaf(e1, e1, p1); p2 = fmaf(e2, e2, p2); p3 = fmaf(e3, e3, p3);
              }
              for (; j < d_full; j += step) {
                const float e0 = to_float(row[HAS_IDX ? tab[j] : j]);
                p0 = fmaf(e0, e0, p0);
              }
              acc = ((double)p0 + (double)p1) + ((double)p2 + (double)p3);
            } else {
              double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
              for (; j + 3 * step < d_full; j += 4 * step) {
                a0 += sq_as_double<T>(row[HAS_IDX ? tab[j] : j]);
                a1 += sq_as_double<T>(row[HAS_IDX ? tab[j + step] : j + step]);
                a2 += sq_as_double<T>(row[HAS_IDX ? tab[j + 2 * step] : j + 2 * step]);
                a3 += sq_as_double<T>(row[HAS_IDX ? tab[j + 3 * step] : j + 3 * step]);
              }
              for (; j < d_full; j += step) a0 += sq_as_double<T>(row[HAS_IDX ? tab[j] : j]);
              acc = (a0 + a1) + (a2 + a3);
            }
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const double other = __shfl_xor_sync(0xffffffffu, acc, o);
          if (o < L) acc += other;
        }
        if (v < rows && sl == 0 && grp / split < slots / split) parts[v * kMaxSplit + part] = acc;
      }
      __syncthreads();
      if (tid < rows) {
        double acc = 0.0;
        for (int p = 0; p < split; ++p) acc += parts[tid * kMaxSplit + p];
        double r = sqrt(acc);
        r = (r == 0.0) ? 1.0 : r;
        nrm[tid] = r;
        if constexpr (sizeof(T) == 4) {
          const float rf = (float)r;  // numpy divides by the float32 norm
          const bool ok = rf >= 1e-30f && rf <= 1e30f;
          rcp[tid] = ok ? (double)(1.0f / rf) : 0.0;
        } else {
          const bool ok = r >= 1e-280 && r <= 1e280;
          rcp[tid] = ok ? 1.0 / r : 0.0;
        }
      }
      __syncthreads();
    }

    // phase 2: gather (+ scale) and stream out.  A thread owns one 16-byte column chunk (gather
    // indices hoisted into registers) and walks the tile's vectors; the chunks left over after the
    // last full round of 256 are shared out by ROWS between thread groups, so short rows (e.g. 288
    // chunks) keep every warp busy
    T *dst = out + (size_t)t * tv * d_full;
    if (vec_ok) {
      const int nchunk = d_full / V;  // d_full % V == 0 when vec_ok
      auto sweep = [&](int c, int v_first, auto step_tag, int v_step_rt) {
        // full rounds walk every vector (compile-time stride 1, so the 4-way unroll really happens)
        const int v_step = decltype(step_tag)::value ? 1 : v_step_rt;
        int col[V];
#pragma unroll
        for (int k = 0; k < V; ++k) col[k] = HAS_IDX ? tab[c * V + k] : (c * V + k);
#pragma unroll 4
        for (int v = v_first; v < rows; v += v_step) {
          const T *row = sm + (size_t)v * in_stride;
          T e[V];
          const double nv = NORM ? nrm[v] : 1.0, rv = NORM ? rcp[v] : 1.0;
#pragma unroll
          for (int k = 0; k < V; ++k) e[k] = NORM ? scale_elem<T>(row[col[k]], nv, rv) : row[col[k]];
          int4 q;
          if constexpr (sizeof(T) == 8) {
            q = make_int4((int)(pun64(e[0]) & 0xffffffffu), (int)(pun64(e[0]) >> 32), (int)(pun64(e[1]) & 0xffffffffu),
                          (int)(pun64(e[1]) >> 32));
          } else {
            q = make_int4((int)pun32(e[0]), (int)pun32(e[1]), (int)pun32(e[2]), (int)pun32(e[3]));
          }
          stg_stream16(dst + (size_t)v * d_full + (size_t)c * V, q);
        }
      };
      const int full = (nchunk / kExpandThreads) * kExpandThreads;
      for (int c = tid; c < full; c += kExpandThreads) sweep(c, 0, std::true_type{}, 1);
      const int rem = nchunk - full;
      if (rem > 0) {
        const int rem_pad = (rem + 31) & ~31;
        const int groups = kExpandThreads / rem_pad;  // >= 1
        const int grp = tid / rem_pad, cl = tid - grp * rem_pad;
        if (grp < groups && cl < rem) sweep(full + cl, grp, std::false_type{}, groups);
      }
    } else {
      const int total = rows * d_full;
      for (int q = tid; q < total; q += kExpandThreads) {
        const int v = q / d_full, j = q - v * d_full;
        const T val = sm[(size_t)v * in_stride + (HAS_IDX ? tab[j] : j)];
        dst[q] = NORM ? scale_elem<T>(val, nrm[v], rcp[v]) : val;
      }
    }
    __syncthreads();  // everyone is done with stage s (and nrm) before it is refilled
  }
}

template <typename T, bool NORM>
static int launch_expand(const void *in, void *out, const int32_t *idx, int64_t nvec,
                         int in_stride, int d_full, cudaStream_t stream) {
  if (nvec == 0) return SOAP_OK;
  const size_t row_bytes = (size_t)in_stride * sizeof(T);
  const char *tile_env = getenv("SOAP_EX_TILE");  // experiment: bytes per stage
  // ~40 KB per stage (two to four CTAs per SM); the fp32 expand+normalise kernel also carries the table and
  // the slot weights and runs best with three CTAs of 32 KB stages (A/B: 0.84 -> 0.89 of the HBM peak)
  const int tile_bytes = tile_env ? atoi(tile_env) : ((NORM && idx && sizeof(T) == 4) ? 32768 : 40960);
  int tv = (int)(tile_bytes / row_bytes);
  if (tv < 1) tv = 1;
  if (tv > 64) tv = 64;
  // rows that are not 16-byte multiples (fp32 quippy rows: 1198 floats; odd dscribe rows): a tile of 2 or 4 rows is
  // one, so tiles can still be staged by bulk copies -- tv becomes a multiple of that row count
  const int row_mult = row_bytes % 16 == 0 ? 1 : (row_bytes % 8 == 0 ? 2 : (row_bytes % 4 == 0 ? 4 : 0));
  const int d_pad = (d_full + 3) & ~3;
  auto smem_for = [&](int rows) {
    const size_t stage_bytes = (((size_t)rows * row_bytes) + 127) & ~(size_t)127;
    return 2 * stage_bytes + (idx ? (size_t)d_pad * 4 : 0) + ((NORM && idx && sizeof(T) == 4) ? (size_t)((in_stride + 3) & ~3) * sizeof(T) : 0) +
           2 * ((rows + 1) & ~1) * 8 + (size_t)rows * kMaxSplit * 8 + 16;
  };
  if (row_mult > 1) {
    const int tv_al = tv >= row_mult ? tv / row_mult * row_mult : row_mult;
    if (smem_for(tv_al) <= (size_t)max_smem_optin()) tv = tv_al;  // (very long rows: keep one row per tile, plain copies)
  }
  if ((int64_t)tv > nvec) tv = (int)nvec;
  const size_t smem = smem_for(tv);
  if (smem > (size_t)max_smem_optin())
    return set_error(SOAP_EUNSUPPORTED, "soap_expand: a single vector of %zu bytes does not fit in shared memory",
                     row_bytes);
  const bool in_aligned = (reinterpret_cast<uintptr_t>(in) & 15) == 0;
  const int bulk_ok = ((in_aligned && row_bytes % 16 == 0) ? 1 : 0) |
                      ((in_aligned && row_mult > 0 && ((size_t)tv * row_bytes) % 16 == 0) ? 2 : 0);
  const int vec_ok = ((reinterpret_cast<uintptr_t>(out) & 15) == 0) &&
                     (((size_t)d_full * sizeof(T)) % 16 == 0);
  const int64_t ntiles = (nvec + tv - 1) / tv;
  auto kern = idx ? expand_kernel<T, NORM, true> : expand_kernel<T, NORM, false>;
  SOAP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;  // persistent grid: every CTA that is resident at once, no more
  SOAP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kExpandThreads, smem));
  per_sm = per_sm < 1 ? 1 : per_sm;
  int grid = (int)(ntiles < (int64_t)num_sms() * per_sm ? ntiles : (int64_t)num_sms() * per_sm);
  ProfScope prof(SOAP_PROF_EXPAND, stream);
  kern<<<grid, kExpandThreads, smem, stream>>>(static_cast<const T *>(in), static_cast<T *>(out), idx,
                                                (long long)nvec, in_stride, d_full, tv, bulk_ok, vec_ok);
  return check_launch("expand_kernel");
}

}  // namespace soapb

using namespace soapb;

extern "C" int soap_expand(const void *in, void *out, const int32_t *idx, int64_t nvec,
                           int in_row_stride, int d_full, int elem_size, void *stream) {
  SOAP_REQUIRE(nvec >= 0 && in_row_stride > 0 && d_full > 0, "soap_expand: bad shape");
  SOAP_REQUIRE(nvec == 0 || (in && out), "soap_expand: null buffer");
  SOAP_REQUIRE(idx != nullptr, "soap_expand: idx table is required");
  auto s = static_cast<cudaStream_t>(stream);
  if (elem_size == 8) return launch_expand<uint64_t, false>(in, out, idx, nvec, in_row_stride, d_full, s);
  if (elem_size == 4) return launch_expand<uint32_t, false>(in, out, idx, nvec, in_row_stride, d_full, s);
  return set_error(SOAP_EUNSUPPORTED, "soap_expand: element size %d (only 4 and 8 bytes)", elem_size);
}

extern "C" int soap_expand_normalize(const void *in, void *out, const int32_t *idx, int64_t nvec,
                                     int in_row_stride, int d_full, int dtype, void *stream) {
  SOAP_REQUIRE(nvec >= 0 && in_row_stride > 0 && d_full > 0, "soap_expand_normalize: bad shape");
  SOAP_REQUIRE(nvec == 0 || (in && out), "soap_expand_normalize: null buffer");
  SOAP_REQUIRE(idx != nullptr || d_full == in_row_stride,
               "soap_expand_normalize: identity table needs d_full == in_row_stride");
  auto s = static_cast<cudaStream_t>(stream);
  if (dtype == SOAP_F64) return launch_expand<double, true>(in, out, idx, nvec, in_row_stride, d_full, s);
  if (dtype == SOAP_F32) return launch_expand<float, true>(in, out, idx, nvec, in_row_stride, d_full, s);
  return set_error(SOAP_EUNSUPPORTED, "soap_expand_normalize: dtype %d", dtype);
}
