// pairs.cu -- distance matrix out[i][j] = dist(data[i], spectra[j]) on CUDA cores, with the
// optional fused row-wise argmin / min of applyClassification, plus the dispatcher of soap_pairs.
//
// Reference semantics: getDistanceBetween (src/SOAPify/classify.py:96-117: double loop over a
// scalar callable, output dtype = data.dtype) and applyClassification's argmin / amin over the
// references (classify.py:274-275).  The distance is one of src/SOAPify/distances.py.
//
// This SIMT path serves every shape and every distance kind (row norms are accumulated in the
// same pass, so nothing is pre-normalised or staged): it is the path of the classification
// row (few references: memory-bound GEMV-like) and of small matrices.  Large all-pairs matrices
// go to the tensor-core kernels: pairs_i8.cu (tcgen05 INT8, exact digit-split accumulation, both
// dtypes) by default; pairs_dmma.cu (fp64 DMMA) and pairs_tf32.cu (3xTF32) as the native-format paths.
#include "common.cuh"

namespace soapb {

int pairs_dmma(const double *A, const double *B, double *out, int64_t M, int64_t N, int D, int64_t ldo,
               int kind, int power, void *scratch, cudaStream_t s);
size_t pairs_dmma_scratch(int64_t M, int64_t N, int D);
int pairs_tf32(const float *A, const float *B, float *out, int64_t M, int64_t N, int D, int64_t ldo, int kind,
               int power, void *scratch, cudaStream_t s);
size_t pairs_tf32_scratch(int64_t M, int64_t N, int D);
int pairs_i8_f32(const float *A, const float *B, float *out, int64_t M, int64_t N, int D, int64_t ldo, int kind,
                 int power, void *scratch, cudaStream_t s, int32_t *argmin_out = nullptr, double *min_out = nullptr,
                 long long excl = 0);
int pairs_i8_f64(const double *A, const double *B, double *out, int64_t M, int64_t N, int D, int64_t ldo, int kind,
                 int power, void *scratch, cudaStream_t s, int32_t *argmin_out = nullptr, double *min_out = nullptr,
                 long long excl = 0);
size_t pairs_i8_scratch(int64_t M, int64_t N, int D, int dtype);
size_t pairs_i8_rowmin_scratch(int64_t M, int64_t N, int D, int dtype);

constexpr int kPT = 64;   // tile edge
constexpr int kPK = 16;   // k-slab
constexpr int kPThreads = 256;

// numpy.argmin / numpy.amin order: NaN beats everything, then smaller value, then smaller index
__device__ __forceinline__ bool better(double v, int j, double bv, int bj) {
  const bool vn = v != v, bn = bv != bv;
  if (vn || bn) return vn && (!bn || j < bj);
  return v < bv || (v == bv && j < bj);
}

template <typename T, bool ARGMIN>
__global__ void __launch_bounds__(kPThreads)
    pairs_simt_kernel(const T *__restrict__ A, const T *__restrict__ B, T *__restrict__ out, long long M,
                      long long N, int D, long long ldo, int kind, int power, int *__restrict__ argmin_out,
                      double *__restrict__ min_out) {
  __shared__ double As[kPK][kPT + 4];
  __shared__ double Bs[kPK][kPT + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, 4 x 4 outputs each
  const long long i0 = (long long)blockIdx.x * kPT;
  const long long jt_begin = ARGMIN ? 0 : blockIdx.y;
  const long long jt_end = ARGMIN ? (N + kPT - 1) / kPT : blockIdx.y + 1;

  double best_v[4];
  int best_j[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) { best_v[r] = 0.0; best_j[r] = -1; }

  // loader mapping: thread -> (row lr, 4 consecutive k)
  const int lr = tid >> 2, lk = (tid & 3) * 4;

  for (long long jt = jt_begin; jt < jt_end; ++jt) {
    const long long j0 = jt * kPT;
    double acc[4][4], uu[4], vv[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      uu[r] = vv[r] = 0.0;
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.0;
    }
    for (int k0 = 0; k0 < D; k0 += kPK) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int k = k0 + lk + e;
        const long long ia = i0 + lr, jb = j0 + lr;
        As[lk + e][lr] = (ia < M && k < D) ? (double)A[(size_t)ia * D + k] : 0.0;
        Bs[lk + e][lr] = (jb < N && k < D) ? (double)B[(size_t)jb * D + k] : 0.0;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < kPK; ++k) {
        double a[4], b[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) a[r] = As[k][ty * 4 + r];
#pragma unroll
        for (int c = 0; c < 4; ++c) b[c] = Bs[k][tx * 4 + c];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          uu[r] = fma(a[r], a[r], uu[r]);
          vv[r] = fma(b[r], b[r], vv[r]);
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[r][c] = fma(a[r], b[c], acc[r][c]);
        }
      }
      __syncthreads();
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const long long i = i0 + ty * 4 + r;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const long long j = j0 + tx * 4 + c;
        if (i < M && j < N) {
          double d;
          if (kind == SOAP_KIND_EUCLID) {
            // |x-y|^2 = xx + yy - 2xy would cancel: not offered on this path
            d = sqrt(fmax(uu[r] + vv[c] - 2.0 * acc[r][c], 0.0));
          } else {
            d = soap_epilogue(acc[r][c], uu[r], vv[c], kind, power);
          }
          if (out) out[(size_t)i * ldo + j] = (T)d;
          if (ARGMIN) {
            // the reference takes argmin over the matrix stored as data.dtype
            const double ds = (double)(T)d;
            if (best_j[r] < 0 || better(ds, (int)j, best_v[r], best_j[r])) { best_v[r] = ds; best_j[r] = (int)j; }
          }
        }
      }
    }
  }
  if (ARGMIN) {
    __shared__ double rv[kPT][16];
    __shared__ int rj[kPT][16];
#pragma unroll
    for (int r = 0; r < 4; ++r) { rv[ty * 4 + r][tx] = best_v[r]; rj[ty * 4 + r][tx] = best_j[r]; }
    __syncthreads();
    if (tid < kPT) {
      const long long i = i0 + tid;
      if (i < M) {
        double bv = 0.0;
        int bj = -1;
        for (int x = 0; x < 16; ++x) {
          const int j = rj[tid][x];
          if (j >= 0 && (bj < 0 || better(rv[tid][x], j, bv, bj))) { bv = rv[tid][x]; bj = j; }
        }
        if (argmin_out) argmin_out[i] = bj;
        if (min_out) min_out[i] = bv;
      }
    }
  }
}

template <typename T>
static int launch_pairs_simt(const void *A, const void *B, void *out, int64_t M, int64_t N, int D, int64_t ldo,
                             int kind, int power, int32_t *argmin_out, double *min_out, cudaStream_t s) {
  const bool am = argmin_out || min_out;
  dim3 grid((unsigned)((M + kPT - 1) / kPT), am ? 1u : (unsigned)((N + kPT - 1) / kPT));
  SOAP_REQUIRE(grid.y <= 65535, "soap_pairs: n_spectra too large for the SIMT path");
  if (am)
    pairs_simt_kernel<T, true><<<grid, kPThreads, 0, s>>>(static_cast<const T *>(A), static_cast<const T *>(B),
                                                          static_cast<T *>(out), M, N, D, ldo, kind, power,
                                                          argmin_out, min_out);
  else
    pairs_simt_kernel<T, false><<<grid, kPThreads, 0, s>>>(static_cast<const T *>(A), static_cast<const T *>(B),
                                                           static_cast<T *>(out), M, N, D, ldo, kind, power,
                                                           nullptr, nullptr);
  return check_launch("pairs_simt_kernel");
}

static bool tensor_eligible(int64_t M, int64_t N, int D, int kind, int dtype, int path) {
  if (kind == SOAP_KIND_EUCLID) return false;
  if (path == SOAP_PAIRS_TENSOR_NATIVE && dtype == SOAP_F64) return D % 2 == 0 && M >= 128 && N >= 128;
  return M >= 128 && N >= 256;
}

}  // namespace soapb

using namespace soapb;

extern "C" size_t soap_pairs_scratch_bytes(int64_t n_data, int64_t n_spectra, int dim, int dtype, int path) {
  if (path == SOAP_PAIRS_SIMT) return 0;
  if (path == SOAP_PAIRS_TENSOR_NATIVE)
    return dtype == SOAP_F64 ? pairs_dmma_scratch(n_data, n_spectra, dim) : pairs_tf32_scratch(n_data, n_spectra, dim);
  return pairs_i8_scratch(n_data, n_spectra, dim, dtype);
}

extern "C" size_t soap_pairs_rowmin_scratch_bytes(int64_t n_data, int64_t n_spectra, int dim, int dtype) {
  return pairs_i8_rowmin_scratch(n_data, n_spectra, dim, dtype);
}

// row minima / arg-minima on the tensor-core path (the matrix itself optional)
static int pairs_rowmin_tensor(const void *data, const void *spectra, void *out, int64_t n_data, int64_t n_spectra, int dim,
                               int64_t ldo, int kind, int power, int dtype, long long excl, int32_t *argmin_out,
                               double *min_out, void *scratch, cudaStream_t s) {
  SOAP_REQUIRE(kind == SOAP_KIND_SIMPLE || kind == SOAP_KIND_KERNEL || kind == SOAP_KIND_NORMALIZED,
               "soap_pairs: row minima on the tensor-core path exist for the three SOAP distances only (kind %d)", kind);
  SOAP_REQUIRE(n_spectra < 2147483647LL, "soap_pairs: too many columns for an int32 arg-minimum");
  SOAP_REQUIRE(scratch != nullptr, "soap_pairs: row minima on the tensor-core path need soap_pairs_rowmin_scratch_bytes() of scratch");
  if (dtype == SOAP_F64)
    return pairs_i8_f64(static_cast<const double *>(data), static_cast<const double *>(spectra), static_cast<double *>(out),
                        n_data, n_spectra, dim, ldo, kind, power, scratch, s, argmin_out, min_out, excl);
  return pairs_i8_f32(static_cast<const float *>(data), static_cast<const float *>(spectra), static_cast<float *>(out), n_data,
                      n_spectra, dim, ldo, kind, power, scratch, s, argmin_out, min_out, excl);
}

extern "C" int soap_pairs_rowmin(const void *data, const void *spectra, void *out, int64_t n_data, int64_t n_spectra, int dim,
                                 int64_t ldo, int kind, int power, int dtype, int64_t exclude_offset, int32_t *argmin_out,
                                 double *min_out, void *scratch, void *stream) {
  SOAP_REQUIRE(n_data >= 0 && n_spectra >= 0 && dim > 0, "soap_pairs_rowmin: bad shape");
  if (n_data == 0 || n_spectra == 0) return SOAP_OK;
  SOAP_REQUIRE(data && spectra && (argmin_out || min_out), "soap_pairs_rowmin: null buffer");
  SOAP_REQUIRE(!out || ldo >= n_spectra, "soap_pairs_rowmin: ldo %lld < n_spectra", (long long)ldo);
  SOAP_REQUIRE(dtype == SOAP_F32 || dtype == SOAP_F64, "soap_pairs_rowmin: dtype %d", dtype);
  SOAP_REQUIRE(tensor_eligible(n_data, n_spectra, dim, kind, dtype, SOAP_PAIRS_TENSOR),
               "soap_pairs_rowmin: shape not eligible for the tensor-core path (>= 128 x 256)");
  return pairs_rowmin_tensor(data, spectra, out, n_data, n_spectra, dim, ldo, kind, power, dtype, (long long)exclude_offset,
                             argmin_out, min_out, scratch, static_cast<cudaStream_t>(stream));
}

extern "C" int soap_pairs(const void *data, const void *spectra, void *out, int64_t n_data, int64_t n_spectra,
                          int dim, int64_t ldo, int kind, int power, int dtype, int path, int32_t *argmin_out,
                          double *min_out, void *scratch, void *stream) {
  SOAP_REQUIRE(n_data >= 0 && n_spectra >= 0 && dim > 0, "soap_pairs: bad shape");
  if (n_data == 0 || n_spectra == 0) return SOAP_OK;
  SOAP_REQUIRE(data && spectra, "soap_pairs: null input");
  SOAP_REQUIRE(out || argmin_out || min_out, "soap_pairs: no output requested");
  SOAP_REQUIRE(!out || ldo >= n_spectra, "soap_pairs: ldo %lld < n_spectra", (long long)ldo);
  SOAP_REQUIRE(dtype == SOAP_F32 || dtype == SOAP_F64, "soap_pairs: dtype %d", dtype);
  SOAP_REQUIRE(kind >= SOAP_KIND_SIMPLE && kind <= SOAP_KIND_NORMALIZED_FUSED, "soap_pairs: kind %d", kind);
  auto s = static_cast<cudaStream_t>(stream);
  const bool am = argmin_out || min_out;
  bool tensor = false;
  if (path == SOAP_PAIRS_TENSOR || path == SOAP_PAIRS_TENSOR_NATIVE) {
    SOAP_REQUIRE(!am || path == SOAP_PAIRS_TENSOR, "soap_pairs: argmin/min is fused on the SIMT and the int8 tensor path only");
    SOAP_REQUIRE(tensor_eligible(n_data, n_spectra, dim, kind, dtype, path),
                 "soap_pairs: shape/kind not eligible for the tensor-core path (>= 128 x 256; native fp64: even D, "
                 ">= 128 x 128; no EUCLID)");
    tensor = true;
  } else if (path == SOAP_PAIRS_AUTO) {
    // (rows shorter than 32 elements stay on the SIMT kernel: nothing to gain from a K block that is mostly padding,
    // and the per-row fixed-point digits of the int8 path are at the edge of the 1e-6 rule for such short float32 rows)
    tensor = !am && scratch != nullptr && dim >= 32 && tensor_eligible(n_data, n_spectra, dim, kind, dtype, SOAP_PAIRS_TENSOR) &&
             (double)n_data * (double)n_spectra >= 1.0e6;
  }
  if (tensor && path == SOAP_PAIRS_AUTO) path = SOAP_PAIRS_TENSOR;
  if (tensor) {
    SOAP_REQUIRE(scratch != nullptr, "soap_pairs: the tensor-core path needs soap_pairs_scratch_bytes() of scratch");
    if (path == SOAP_PAIRS_TENSOR_NATIVE) {
      if (dtype == SOAP_F64)
        return pairs_dmma(static_cast<const double *>(data), static_cast<const double *>(spectra),
                          static_cast<double *>(out), n_data, n_spectra, dim, ldo, kind, power, scratch, s);
      return pairs_tf32(static_cast<const float *>(data), static_cast<const float *>(spectra),
                        static_cast<float *>(out), n_data, n_spectra, dim, ldo, kind, power, scratch, s);
    }
    if (am)  // scratch must then hold soap_pairs_rowmin_scratch_bytes(); no column is excluded
      return pairs_rowmin_tensor(data, spectra, out, n_data, n_spectra, dim, ldo, kind, power, dtype, SOAP_NO_EXCLUDE,
                                 argmin_out, min_out, scratch, s);
    if (dtype == SOAP_F64)
      return pairs_i8_f64(static_cast<const double *>(data), static_cast<const double *>(spectra),
                          static_cast<double *>(out), n_data, n_spectra, dim, ldo, kind, power, scratch, s);
    return pairs_i8_f32(static_cast<const float *>(data), static_cast<const float *>(spectra),
                        static_cast<float *>(out), n_data, n_spectra, dim, ldo, kind, power, scratch, s);
  }
  if (dtype == SOAP_F64)
    return launch_pairs_simt<double>(data, spectra, out, n_data, n_spectra, dim, ldo, kind, power, argmin_out,
                                     min_out, s);
  return launch_pairs_simt<float>(data, spectra, out, n_data, n_spectra, dim, ldo, kind, power, argmin_out,
                                  min_out, s);
}
