// classify.cu -- f1: distances from every (frame, atom) vector of a COMPRESSED SOAP dataset to a small
// dictionary of reference spectra, with the fused arg-minimum of applyClassification.
//
// Reference semantics: getDistancesFromRef (src/SOAPify/classify.py:120-160: expand with the
// mono-species table when the dataset is compressed, optionally normalise, then
// getDistanceBetween(frame, references.spectra, f)) and applyClassification's argmin / amin over the
// references (classify.py:274-275).
//
// Memory-bound GEMV-like work (K references, K small): the expansion is folded into the references --
// with R'[k][i] = sum over the expanded slots j that gather compressed slot i of r_k[j], the dot of
// the EXPANDED vector with r_k is x . R'_k -- and the norm of the expanded vector is the
// multiplicity-weighted norm of the compressed one.  The kernel therefore reads Dc elements per
// vector and nothing else: no expanded / normalised copy of the dataset ever exists.
// One warp per vector, 128-bit streaming loads, R' and the multiplicities in shared memory,
// K+1 warp reductions per vector, numpy's argmin rules (first NaN, then first minimum).
#include "tc05.cuh"

namespace soapb {

constexpr int kClsThreads = 256;
constexpr int kClsTK = 8;  // references handled per pass over a vector

__device__ __forceinline__ bool cls_better(double v, int j, double bv, int bj) {
  const bool vn = v != v, bn = bv != bv;
  if (vn || bn) return vn && (!bn || j < bj);
  return v < bv || (v == bv && j < bj);
}

template <typename T>
__global__ void __launch_bounds__(kClsThreads, 2)
    classify_kernel(const T *__restrict__ X, const double *__restrict__ mult, const double *__restrict__ refw,
                    const double *__restrict__ ref_norm2, long long nvec, int D, int K, int kind, int power,
                    int normalize_x, double *__restrict__ dist_out, int *__restrict__ argmin_out,
                    double *__restrict__ min_out) {
  extern __shared__ __align__(16) double smem_cls[];
  double *sm_m = smem_cls;             // [D] multiplicities (1.0 when mult == NULL)
  double *sm_r = smem_cls + D;         // [K][D] folded references
  for (int i = threadIdx.x; i < D; i += kClsThreads) sm_m[i] = mult ? mult[i] : 1.0;
  for (int i = threadIdx.x; i < K * D; i += kClsThreads) sm_r[i] = refw[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long nwarps = (long long)gridDim.x * (kClsThreads / 32);
  for (long long v = (long long)blockIdx.x * (kClsThreads / 32) + warp; v < nvec; v += nwarps) {
    const T *x = X + (size_t)v * D;
    double best_v = 0.0;
    int best_j = -1;
    double uu = 0.0;
    for (int k0 = 0; k0 < K; k0 += kClsTK) {
      const int kn = min(kClsTK, K - k0);
      double acc[kClsTK], nn = 0.0;
#pragma unroll
      for (int k = 0; k < kClsTK; ++k) acc[k] = 0.0;
      // four independent loads in flight per lane (the second and later passes hit L1/L2)
      int i = lane;
      for (; i + 96 < D; i += 128) {
        double xv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) xv[u] = (double)x[i + 32 * u];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int e = i + 32 * u;
          nn = fma(xv[u] * sm_m[e], xv[u], nn);
#pragma unroll
          for (int k = 0; k < kClsTK; ++k)
            if (k < kn) acc[k] = fma(xv[u], sm_r[(size_t)(k0 + k) * D + e], acc[k]);
        }
      }
      for (; i < D; i += 32) {
        const double xi = (double)x[i];
        nn = fma(xi * sm_m[i], xi, nn);
#pragma unroll
        for (int k = 0; k < kClsTK; ++k)
          if (k < kn) acc[k] = fma(xi, sm_r[(size_t)(k0 + k) * D + i], acc[k]);
      }
      if (k0 == 0) uu = warp_sum(nn);
#pragma unroll
      for (int k = 0; k < kClsTK; ++k) {
        if (k >= kn) break;
        const double uv = warp_sum(acc[k]);
        const int j = k0 + k;
        double d;
        if (kind == SOAP_KIND_NORMALIZED) {
          // SOAPdistanceNormalized on (optionally normalised) data: the reference spectra are used as given
          const double nu = normalize_x ? ((uu == 0.0) ? 1.0 : sqrt(uu)) : 1.0;
          d = sqrt(fabs(2.0 - 2.0 * (uv / nu)));
        } else {
          d = soap_epilogue(uv, uu, ref_norm2[j], kind, power);  // scale invariant in x
        }
        // the reference stores each frame's distances as data.dtype (classify.py:113) before amin / argmin
        if constexpr (sizeof(T) == 4) d = (double)(float)d;
        if (dist_out && lane == 0) dist_out[(size_t)v * K + j] = d;
        if (best_j < 0 || cls_better(d, j, best_v, best_j)) { best_v = d; best_j = j; }
      }
    }
    if (lane == 0) {
      if (argmin_out) argmin_out[v] = best_j;
      if (min_out) min_out[v] = best_v;
    }
  }
}

// ---- tiled variant: LANE = VECTOR ---------------------------------------------------------------
// A CTA takes 32 consecutive vectors at a time.  Their rows (or a piece of them when the rows are
// long) are dropped by the TMA engine into shared-memory slots 16*odd bytes apart, so that lane l
// walks ITS OWN vector with conflict-free 128-bit loads while the folded references and the
// multiplicities are read as warp-wide BROADCASTS (one wavefront for all 32 vectors).  No shuffle
// reduction exists: a lane ends with the complete sums of its vector.  8 consumer warps split the
// 16-byte units, warp 8 is the TMA producer, warp 9 sums the consumers' partials of a tile, applies
// the distance and the arg-minimum and writes the results; tiles and partial-sum buffers are double
// buffered and every hand-off is an mbarrier (no __syncthreads after the prologue).
constexpr int kCtWarps = 8;
constexpr int kCtThreads = (kCtWarps + 3) * 32;  // + TMA producer warp + two finisher warps (alternating tiles)
constexpr int kCtMaxK = 32;

// VPL = vectors per lane: a tile holds 32 * VPL vectors and lane l owns vectors l, l + 32, ...  One broadcast
// load of the references / multiplicities then serves VPL vectors, which takes the kernel off the
// shared-memory roofline for small dictionaries (K <= 8: 13 -> 8.5 wavefronts per vector unit).
template <typename T, int KMAX, int VPL>
__global__ void __launch_bounds__(kCtThreads, 1)
    classify_tiled_kernel(const T *__restrict__ X, const double *__restrict__ mult, const double *__restrict__ refw,
                          const double *__restrict__ ref_norm2, long long nvec, int D, int K, int kind, int power,
                          int normalize_x, double *__restrict__ dist_out, int *__restrict__ argmin_out,
                          double *__restrict__ min_out, int upp, int npieces, int stride, int nstage,
                          const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) unsigned char smem_ct[];
  constexpr int UE = 16 / (int)sizeof(T);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int Dp = npieces * upp * UE;                   // padded dimension (multiple of the unit)
  constexpr int TV = 32 * VPL;                          // vectors per tile
  unsigned char *stage0 = smem_ct;                     // [nstage][TV slots][stride]
  double *sm_m = reinterpret_cast<double *>(smem_ct + (size_t)nstage * TV * stride);  // [Dp]
  double *sm_r = sm_m + Dp;                            // [K][Dp]
  double *red = sm_r + (size_t)K * Dp;                 // [2][kCtWarps][K+1][TV]
  const size_t red_elems = (size_t)kCtWarps * (K + 1) * TV;
  uint64_t *full = reinterpret_cast<uint64_t *>(red + 2 * red_elems);  // [nstage]
  uint64_t *empty = full + nstage, *red_full = empty + nstage, *red_empty = red_full + 2;

  for (int i = tid; i < Dp; i += kCtThreads) sm_m[i] = i < D ? (mult ? mult[i] : 1.0) : 0.0;
  for (int i = tid; i < K * Dp; i += kCtThreads) {
    const int k = i / Dp, e = i - k * Dp;
    sm_r[i] = e < D ? refw[(size_t)k * D + e] : 0.0;
  }
  if (tid == 0) {
    for (int s = 0; s < nstage; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], kCtWarps);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&red_full[s], kCtWarps);
      mbar_init(&red_empty[s], 1);
    }
    fence_mbar_init();
  }
  __syncthreads();

  const long long ntiles = (nvec + TV - 1) / TV;
  const long long my_first = blockIdx.x;
  if (warp == kCtWarps) {
    // ===== producer: ONE tensor-map box (32 vectors x one piece) per stage ==========================
    // rows beyond nvec and columns beyond D are zero-filled and still counted in the transaction
    int s = 0, ph = 0;
    long long it = 0;
    for (long long t = my_first; t < ntiles; t += gridDim.x) {
      for (int p = 0; p < npieces; ++p, ++it) {
        if (it >= nstage) mbar_wait(&empty[s], (uint32_t)(ph ^ 1));
        if (lane == 0) {
          mbar_expect_tx(&full[s], (uint32_t)TV * (uint32_t)stride);
          tma_load_2d(stage0 + (size_t)s * TV * stride, &tmap, p * (stride / 8), (int)(t * TV), &full[s]);
        }
        if (++s == nstage) s = 0, ph ^= 1;
      }
    }
  } else if (warp > kCtWarps) {
    // ===== finishers: sums of the consumers' partials -> distances, arg-minimum, results ==========
    // (the fp64 sqrt / divide chains of K distances take about as long as the consumers need for a
    // tile: two warps take alternate tiles, each owning one of the two partial-sum buffers)
    const int fin = warp - kCtWarps - 1;
    long long tcount = fin;
    for (long long t = my_first + (long long)fin * gridDim.x; t < ntiles; t += 2LL * gridDim.x, tcount += 2) {
      const int rb = (int)(tcount & 1);
      mbar_wait(&red_full[rb], (uint32_t)((tcount >> 1) & 1));
      const double *r = red + rb * red_elems;
#pragma unroll
      for (int vi = 0; vi < VPL; ++vi) {
        const int vl = vi * 32 + lane;
        const long long v = t * TV + vl;
        double uu = 0.0;
        for (int w = 0; w < kCtWarps; ++w) uu += r[((size_t)w * (K + 1)) * TV + vl];
        double best_v = 0.0;
        int best_j = -1;
        const double nu = (kind == SOAP_KIND_NORMALIZED && normalize_x) ? ((uu == 0.0) ? 1.0 : sqrt(uu)) : 1.0;
        double d[KMAX];
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {  // independent chains: the compiler interleaves them
          if (k < K) {
            double uv = 0.0;
#pragma unroll
            for (int w = 0; w < kCtWarps; ++w) uv += r[((size_t)w * (K + 1) + k + 1) * TV + vl];
            d[k] = (kind == SOAP_KIND_NORMALIZED) ? sqrt(fabs(2.0 - 2.0 * (uv / nu)))
                                                  : soap_epilogue(uv, uu, ref_norm2[k], kind, power);
            // the reference stores each frame's distances as data.dtype (classify.py:113) before amin / argmin
            if constexpr (sizeof(T) == 4) d[k] = (double)(float)d[k];
          }
        }
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
          if (k < K) {
            if (v < nvec && dist_out) dist_out[(size_t)v * K + k] = d[k];
            if (best_j < 0 || cls_better(d[k], k, best_v, best_j)) { best_v = d[k]; best_j = k; }
          }
        }
        if (v < nvec) {
          if (argmin_out) argmin_out[v] = best_j;
          if (min_out) min_out[v] = best_v;
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&red_empty[rb]);
    }
  } else {
    // ===== consumers ================================================================================
    const uint32_t st_addr = smem_u32(stage0), m_addr = smem_u32(sm_m), r_addr = smem_u32(sm_r);
    long long tcount = 0;
    int s = 0, ph = 0;
    for (long long t = my_first; t < ntiles; t += gridDim.x) {
      double acc[VPL][KMAX + 1];  // statically indexed everywhere: stays in registers
#pragma unroll
      for (int vi = 0; vi < VPL; ++vi)
#pragma unroll
        for (int q = 0; q <= KMAX; ++q) acc[vi][q] = 0.0;
      for (int p = 0; p < npieces; ++p) {
        const int e0 = p * upp * UE;
        const int punits = (min(D - e0, upp * UE) + UE - 1) / UE;
        mbar_wait(&full[s], (uint32_t)ph);
        const uint32_t xrow = st_addr + (uint32_t)((s * TV + lane) * stride);  // vector vi of this lane: + vi * 32 slots
        for (int u = warp; u < punits; u += kCtWarps) {
          if constexpr (sizeof(T) == 8) {
            double2 xv[VPL];
#pragma unroll
            for (int vi = 0; vi < VPL; ++vi) xv[vi] = lds_d2(xrow + (uint32_t)(vi * 32) * (uint32_t)stride + 16u * u);
            const uint32_t eo = (uint32_t)(e0 + 2 * u) * 8u;
            const double2 mv = lds_d2(m_addr + eo);
#pragma unroll
            for (int vi = 0; vi < VPL; ++vi) {
              acc[vi][0] = fma(xv[vi].x * mv.x, xv[vi].x, acc[vi][0]);
              acc[vi][0] = fma(xv[vi].y * mv.y, xv[vi].y, acc[vi][0]);
            }
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
              if (k < K) {
                const double2 rv = lds_d2(r_addr + (uint32_t)k * (uint32_t)Dp * 8u + eo);
#pragma unroll
                for (int vi = 0; vi < VPL; ++vi) {
                  acc[vi][k + 1] = fma(xv[vi].x, rv.x, acc[vi][k + 1]);
                  acc[vi][k + 1] = fma(xv[vi].y, rv.y, acc[vi][k + 1]);
                }
              }
            }
          } else {
            double xs[VPL][4];
#pragma unroll
            for (int vi = 0; vi < VPL; ++vi) {
              const float4 xv = lds_f4(xrow + (uint32_t)(vi * 32) * (uint32_t)stride + 16u * u);
              xs[vi][0] = (double)xv.x; xs[vi][1] = (double)xv.y; xs[vi][2] = (double)xv.z; xs[vi][3] = (double)xv.w;
            }
            const uint32_t eo = (uint32_t)(e0 + 4 * u) * 8u;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const double2 mv = lds_d2(m_addr + eo + 16u * h);
#pragma unroll
              for (int vi = 0; vi < VPL; ++vi) {
                acc[vi][0] = fma(xs[vi][2 * h] * mv.x, xs[vi][2 * h], acc[vi][0]);
                acc[vi][0] = fma(xs[vi][2 * h + 1] * mv.y, xs[vi][2 * h + 1], acc[vi][0]);
              }
            }
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
              if (k < K) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const double2 rv = lds_d2(r_addr + (uint32_t)k * (uint32_t)Dp * 8u + eo + 16u * h);
#pragma unroll
                  for (int vi = 0; vi < VPL; ++vi) {
                    acc[vi][k + 1] = fma(xs[vi][2 * h], rv.x, acc[vi][k + 1]);
                    acc[vi][k + 1] = fma(xs[vi][2 * h + 1], rv.y, acc[vi][k + 1]);
                  }
                }
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        if (++s == nstage) s = 0, ph ^= 1;
      }
      // hand the K+1 partial sums of every vector to the finisher warp
      const int rb = (int)(tcount & 1);
      if (tcount >= 2) mbar_wait(&red_empty[rb], (uint32_t)(((tcount >> 1) - 1) & 1));
      double *r = red + rb * red_elems;
#pragma unroll
      for (int vi = 0; vi < VPL; ++vi)
#pragma unroll
        for (int q = 0; q <= KMAX; ++q)
          if (q <= K) r[((size_t)warp * (K + 1) + q) * TV + vi * 32 + lane] = acc[vi][q];
      __syncwarp();
      if (lane == 0) mbar_arrive(&red_full[rb]);
      ++tcount;
    }
  }
}

}  // namespace soapb

using namespace soapb;

extern "C" int soap_classify(const void *X, const void *mult, const double *refw, const double *ref_norm2,
                             int64_t nvec, int dim, int n_refs, int kind, int power, int normalize_x, int dtype,
                             double *dist_out, int32_t *argmin_out, double *min_out, void *stream) {
  SOAP_REQUIRE(nvec >= 0 && dim > 0 && n_refs > 0, "soap_classify: bad shape");
  if (nvec == 0) return SOAP_OK;
  SOAP_REQUIRE(X && refw, "soap_classify: null input");
  SOAP_REQUIRE(dist_out || argmin_out || min_out, "soap_classify: no output requested");
  SOAP_REQUIRE(kind == SOAP_KIND_NORMALIZED || ref_norm2 != nullptr, "soap_classify: reference norms are required for this kind");
  SOAP_REQUIRE(kind == SOAP_KIND_SIMPLE || kind == SOAP_KIND_KERNEL || kind == SOAP_KIND_NORMALIZED,
               "soap_classify: kind %d", kind);
  auto s = static_cast<cudaStream_t>(stream);
  {
    // tiled lane=vector kernel: rows must be 16-byte multiples; references + two tile stages must fit
    const int es = dtype == SOAP_F64 ? 8 : 4, ue = 16 / es;
    const bool aligned = ((reinterpret_cast<uintptr_t>(X) & 15) == 0) && (((size_t)dim * es) % 16 == 0);
    const int n_units = (dim + ue - 1) / ue;
    const size_t budget = (size_t)max_smem_optin() - 2048;
    if (aligned && n_refs <= kCtMaxK && (dtype == SOAP_F64 || dtype == SOAP_F32)) {
      // pieces of an odd number of 16-byte units (slots 16*odd bytes apart, <= 127 units = one box row);
      // as many stages as fit, preferring >= 64 KB in flight per SM and then the fewest pieces.  Small
      // dictionaries take two vectors per lane (64-vector tiles) when that still leaves >= 48 KB in flight.
      auto plan = [&](int vpl, int &o_upp, int &o_pieces, int &o_stage) -> size_t {
        size_t best_score = 0;
        o_upp = 0;
        for (int npieces = 1; npieces <= n_units; ++npieces) {
          const int upp = ((n_units + npieces - 1) / npieces) | 1;
          if (upp > 127 || upp > n_units) continue;
          const int real_pieces = (n_units + upp - 1) / upp;
          const size_t Dp = (size_t)real_pieces * upp * ue;
          const size_t fixed = (size_t)(n_refs + 1) * Dp * 8 + (size_t)2 * kCtWarps * (n_refs + 1) * 32 * vpl * 8 + 256;
          for (int st = 4; st >= 2; --st) {
            if ((size_t)st * 32 * vpl * 16 * upp + fixed > budget) continue;
            size_t score = (size_t)(st - 1) * 32 * vpl * 16 * upp;
            if (score > 65536) score = 65536;
            if (score > best_score) best_score = score, o_upp = upp, o_pieces = real_pieces, o_stage = st;
            break;
          }
        }
        return best_score;
      };
      int vpl = 1, best_upp = 0, best_pieces = 0, best_stage = 0;
      if (n_refs <= 8) {
        int u2, p2, s2;
        if (plan(2, u2, p2, s2) >= 49152) vpl = 2, best_upp = u2, best_pieces = p2, best_stage = s2;
      }
      if (vpl == 1) plan(1, best_upp, best_pieces, best_stage);
      if (best_upp > 0) {
        const int upp = best_upp, real_pieces = best_pieces, stride = 16 * best_upp, nstage = best_stage, tv = 32 * vpl;
        const size_t Dp = (size_t)real_pieces * upp * ue;
        const size_t smem_t = (size_t)nstage * tv * stride + (size_t)(n_refs + 1) * Dp * 8 +
                              (size_t)2 * kCtWarps * (n_refs + 1) * tv * 8 + 256;
        const long long ntiles = (nvec + tv - 1) / tv;
        const int grid = (int)(ntiles < num_sms() ? ntiles : num_sms());
        CUtensorMap tmap;
        int rc = make_tensor_map_2d(&tmap, X, CU_TENSOR_MAP_DATA_TYPE_UINT64, 8, nvec, (int64_t)dim * es / 8, stride / 8, tv,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE);
        if (rc) return rc;
        ProfScope prof(SOAP_PROF_PAIRS, s);
#define CT_LAUNCH(T, KM, VP)                                                                                          \
  do {                                                                                                                \
    SOAP_CUDA(cudaFuncSetAttribute(classify_tiled_kernel<T, KM, VP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t)); \
    classify_tiled_kernel<T, KM, VP><<<grid, kCtThreads, smem_t, s>>>(                                                \
        static_cast<const T *>(X), static_cast<const double *>(mult), refw, ref_norm2, nvec, dim, n_refs, kind, power, \
        normalize_x, dist_out, argmin_out, min_out, upp, real_pieces, stride, nstage, tmap);                          \
  } while (0)
        if (dtype == SOAP_F64) {
          if (n_refs <= 8) { if (vpl == 2) CT_LAUNCH(double, 8, 2); else CT_LAUNCH(double, 8, 1); }
          else if (n_refs <= 16) CT_LAUNCH(double, 16, 1); else CT_LAUNCH(double, 32, 1);
        } else {
          if (n_refs <= 8) { if (vpl == 2) CT_LAUNCH(float, 8, 2); else CT_LAUNCH(float, 8, 1); }
          else if (n_refs <= 16) CT_LAUNCH(float, 16, 1); else CT_LAUNCH(float, 32, 1);
        }
#undef CT_LAUNCH
        return check_launch("classify_tiled_kernel");
      }
    }
  }
  const size_t smem = (size_t)(n_refs + 1) * dim * sizeof(double);
  if (smem > (size_t)max_smem_optin() - 1024)
    return set_error(SOAP_EUNSUPPORTED, "soap_classify: %d references of dimension %d do not fit in shared memory; use soap_pairs",
                     n_refs, dim);
  int per_sm = (int)((size_t)(max_smem_optin() - 2048) / (smem + 1024));  // latency hiding comes from resident warps
  per_sm = per_sm < 1 ? 1 : (per_sm > 6 ? 6 : per_sm);
  const long long want = (nvec + kClsThreads / 32 - 1) / (kClsThreads / 32);
  const int grid = (int)(want < (long long)num_sms() * per_sm ? want : (long long)num_sms() * per_sm);
  ProfScope prof(SOAP_PROF_PAIRS, s);
  if (dtype == SOAP_F64) {
    SOAP_CUDA(cudaFuncSetAttribute(classify_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    classify_kernel<double><<<grid, kClsThreads, smem, s>>>(static_cast<const double *>(X), static_cast<const double *>(mult),
                                                            refw, ref_norm2, nvec, dim, n_refs, kind, power, normalize_x,
                                                            dist_out, argmin_out, min_out);
  } else if (dtype == SOAP_F32) {
    SOAP_CUDA(cudaFuncSetAttribute(classify_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    classify_kernel<float><<<grid, kClsThreads, smem, s>>>(static_cast<const float *>(X), static_cast<const double *>(mult),
                                                           refw, ref_norm2, nvec, dim, n_refs, kind, power, normalize_x,
                                                           dist_out, argmin_out, min_out);
  } else {
    return set_error(SOAP_EUNSUPPORTED, "soap_classify: dtype %d", dtype);
  }
  return check_launch("classify_kernel");
}
