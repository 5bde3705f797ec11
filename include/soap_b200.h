/*
 * soap_b200.h -- C ABI of libsoapb200.so: the SOAP post-processing hot path of
 * GMPavanLab/cpctools (SOAPify) as hand-written sm_100a CUDA kernels.
 *
 * The reference is pure Python (no FFI of its own): the functions below are what a
 * ctypes binding inside SOAPify would call in place of the numpy bodies cited on each
 * entry (file:line relative to the reference tree).  INTEGRATION.md shows that binding.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host; the caller owns all
 *     buffers; nothing is allocated or freed behind the caller's back;
 *   - arrays are C-contiguous (row-major), vectors on the last axis;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); every call
 *     is asynchronous on that stream and returns right after the launch;
 *   - return value: 0 on success, a negative SOAP_E* code otherwise; soap_last_error()
 *     gives the message for the calling thread.  No C++ exception crosses the boundary;
 *   - dtype codes: SOAP_F32 / SOAP_F64 select the arithmetic type of the data; integer data
 *     is only ever gathered (soap_expand takes an element size).
 */
#ifndef SOAP_B200_H
#define SOAP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SOAP_B200_ABI_VERSION 1

enum { SOAP_F32 = 0, SOAP_F64 = 1 };

/* distance kinds: src/SOAPify/distances.py */
enum {
  SOAP_KIND_SIMPLE = 0,     /* simpleSOAPdistance   distances.py:22-35 (scipy cosine, clipped)      */
  SOAP_KIND_KERNEL = 1,     /* SOAPdistance(n)      distances.py:54-68 (no clip, NaN if k > 1)      */
  SOAP_KIND_NORMALIZED = 2, /* SOAPdistanceNormalized distances.py:71-83 (sqrt|2-2 x.y|)            */
  SOAP_KIND_DOT = 3,        /* kernelSoap(n): the kernel value (x.y/|x||y|)^n, distances.py:38-51 */
  SOAP_KIND_COSCLIP = 4,    /* simpleKernelSoap: 1 - clip(1 - x.y/sqrt(xx yy), 0, 2), distances.py:7-19 */
  SOAP_KIND_EUCLID = 5,     /* |x - y|_2 : the per-frame norm of timeSOAPsimple, analysis.py:136    */
  SOAP_KIND_EUCLID_NORMALIZED = 7, /* soap_timelag only: |x/|x| - y/|y||_2 with the norms of the EXPANDED vectors taken from
                                    * the compressed rows and the multiplicities (zero norm -> 1): timeSOAPsimple on
                                    * normalizeArray(fillSOAPVectorFromdscribe(x)), analysis.py:192-200, without materialising it */
  SOAP_KIND_NORMALIZED_FUSED = 6 /* normalizeArray (utils.py:140-153, zero norm -> 1) folded into
                               SOAPdistanceNormalized: sqrt|2 - 2 x.y/(nx ny)| on RAW vectors   */
};

enum {
  SOAP_OK = 0,
  SOAP_EINVAL = -1,   /* bad argument (message says which)                     */
  SOAP_ECUDA = -2,    /* CUDA runtime / driver error (message carries the text) */
  SOAP_EUNSUPPORTED = -3
};

/* library / device ---------------------------------------------------------------- */
int soap_abi_version(void);
const char *soap_last_error(void);
/* number of kernels launched through this library by the calling process so far
 * (bench.py reports the delta over its timed region as "gpu_launches") */
int64_t soap_launch_count(void);

/* profiling hook (bench.py): while enabled, every call brackets its MAIN kernel (the one that moves
 * the data: expand, time-lag walk, tensor-core pair tiles) with CUDA events recorded on the call's
 * own stream.  soap_profile_read waits for the recorded events, writes the elapsed milliseconds of
 * the calls of `kernel_id` (SOAP_PROF_*) since the last enable, oldest first, and returns how
 * many there were (at most `capacity` are written). */
enum { SOAP_PROF_EXPAND = 0, SOAP_PROF_TIMELAG = 1, SOAP_PROF_PAIRS = 2 };
int soap_profile_enable(int on);
int soap_profile_read(int kernel_id, float *ms_out_host, int capacity);

/* ---- expansion: fillSOAPVectorFromdscribe, utils.py:271-324 --------------------------
 * out[v][j] = in[v][idx[j]], v < nvec, j < d_full.  Pure gather, bit-exact, any element type
 * of 4 or 8 bytes (elem_size).  `idx` is the table of utils.py:212-268 (or its composition
 * with the quippy permutation utils.py:112-137 / engine.py:260), int32, device memory.
 * in_row_stride = elements between consecutive input rows (>= max(idx)+1; 1198 for raw quippy).
 */
int soap_expand(const void *in, void *out, const int32_t *idx, int64_t nvec, int in_row_stride,
                int d_full, int elem_size, void *stream);

/* ---- expansion fused with normalizeArray (utils.py:271-324 then utils.py:140-153) -----
 * out[v][j] = in[v][idx[j]] / norm_v, norm_v = sqrt(sum_j in[v][idx[j]]^2), norm_v == 0 -> 1.
 * idx == NULL means identity (d_full == in_row_stride): plain normalizeArray.
 * dtype SOAP_F32 / SOAP_F64 (float types only; integer input is promoted by the caller).
 */
int soap_expand_normalize(const void *in, void *out, const int32_t *idx, int64_t nvec,
                          int in_row_stride, int d_full, int dtype, void *stream);

/* ---- time-lagged per-atom distances: timeSOAP, analysis.py:13-69 ------------------------
 * X[F][A][D] (dtype), optional slot multiplicities mult[D] (same dtype as X, or NULL = all 1):
 * with mult = bincount of the gather table the distances are those of the EXPANDED vectors
 * without materialising them (the dot products of utils.py:271's output are weighted dots of
 * its input).  For each window w_k (k < n_windows <= SOAP_MAX_WINDOWS):
 *     t_k[f - w_k][a] = dist(X[f][a][:], X[f - w_k][a][:]),  w_k <= f < F
 * written to t + t_offsets_host[k] (doubles; each block is [F - w_k][A] row-major, always
 * float64 like analysis.py:53).  All windows are produced by ONE pass over X.
 * `scratch` must hold soap_timelag_scratch_bytes(...) bytes.
 */
#define SOAP_MAX_WINDOWS 4
size_t soap_timelag_scratch_bytes(int64_t n_frames, int64_t n_atoms, int dim, int dtype,
                                  const int *windows_host, int n_windows);
/* name of the sweep kernel the last soap_timelag call of this process launched (bench.py's roofline names it):
 * "timelag_kernel" (lane = frame), "timelag_pair_kernel" (two frames per lane), "timelag_chain_kernel" (lane =
 * units, lag history in registers), "timelag_rb_kernel"; "" before the first call */
const char *soap_timelag_last_kernel(void);
int soap_timelag(const void *X, const void *mult, double *t, const int64_t *t_offsets_host,
                 int64_t n_frames, int64_t n_atoms, int dim, const int *windows_host,
                 int n_windows, int kind, int power, int dtype, void *scratch, void *stream);

/* ---- numpy.diff(t.T, axis=-1): analysis.py:67,141,203 -----------------------------------
 * t[rows][A] -> dt[A][rows-1], dt[a][r] = t[r+1][a] - t[r][a].
 */
int soap_diff_transposed(const double *t, double *dt, int64_t rows, int64_t n_atoms, void *stream);

/* ---- distance matrix: getDistanceBetween, classify.py:96-117 ---------------------------
 * out[i][j] = dist(data[i], spectra[j]); data[n_data][D], spectra[n_spectra][D] (dtype);
 * out has leading dimension ldo (elements) and the type of `dtype` (classify.py:113).
 * path: SOAP_PAIRS_AUTO picks; SOAP_PAIRS_SIMT = CUDA-core tiles (any size);
 * SOAP_PAIRS_TENSOR = tcgen05 INT8 tensor-core GEMM with exact digit-split accumulation (3 int8
 * digits for fp32 data, 6 for fp64: fp32 / fp64 accuracy respectively);
 * SOAP_PAIRS_TENSOR_NATIVE = native-format tensor GEMM (fp64: DMMA; fp32: tcgen05 3xTF32, whose
 * fp32 round-toward-zero accumulation limits it to ~2e-5 on d^2).
 * If argmin_out / min_out are non-NULL the row-wise argmin (int32) / min (double) over j is
 * produced as well (applyClassification, classify.py:274-275) and `out` may be NULL
 * (SOAP_PAIRS_SIMT, SOAP_PAIRS_TENSOR or SOAP_PAIRS_AUTO, which then takes the SIMT kernel).
 */
enum { SOAP_PAIRS_AUTO = 0, SOAP_PAIRS_SIMT = 1, SOAP_PAIRS_TENSOR = 2, SOAP_PAIRS_TENSOR_NATIVE = 3 };
/* With path == SOAP_PAIRS_TENSOR the arg-minimum / minimum are fused into the tcgen05 epilogue as well (the three
 * SOAP distances): every (tile column, epilogue warp) keeps the minimum of its columns per row, a second small kernel
 * folds them with numpy's rules (first NaN, smaller value, smaller index).  `scratch` must then be
 * soap_pairs_rowmin_scratch_bytes() large and `out` may be NULL (nothing of the matrix is written).
 * soap_pairs_rowmin() is that call with one more argument: column j == i + exclude_offset is skipped -- the self
 * distance when data is the row block [r0, r0 + n_data) of spectra (exclude_offset = r0; SOAP_NO_EXCLUDE for none):
 * the nearest OTHER vector of every row of the all-pairs matrix without materialising it.
 */
#define SOAP_NO_EXCLUDE INT64_MIN
size_t soap_pairs_rowmin_scratch_bytes(int64_t n_data, int64_t n_spectra, int dim, int dtype);
int soap_pairs_rowmin(const void *data, const void *spectra, void *out, int64_t n_data, int64_t n_spectra, int dim,
                      int64_t ldo, int kind, int power, int dtype, int64_t exclude_offset, int32_t *argmin_out,
                      double *min_out, void *scratch, void *stream);
size_t soap_pairs_scratch_bytes(int64_t n_data, int64_t n_spectra, int dim, int dtype, int path);
int soap_pairs(const void *data, const void *spectra, void *out, int64_t n_data, int64_t n_spectra,
               int dim, int64_t ldo, int kind, int power, int dtype, int path, int32_t *argmin_out,
               double *min_out, void *scratch, void *stream);

/* ---- classification against a small dictionary: getDistancesFromRef, classify.py:120-160, and
 * applyClassification's argmin / amin, classify.py:274-275, fused and straight from COMPRESSED data.
 * X[nvec][dim] (dtype); mult[dim] doubles = slot multiplicities of the mono-species gather table
 * (NULL = 1: data already expanded); refw[n_refs][dim] doubles = reference spectra FOLDED onto the
 * compressed slots (refw[k][i] = sum of r_k[j] over the expanded slots j that gather slot i; the
 * spectra themselves when nothing is compressed); ref_norm2[n_refs] = |r_k|^2 (needed by SIMPLE /
 * KERNEL).  normalize_x != 0 reproduces doNormalize=True for SOAP_KIND_NORMALIZED (zero norm -> 1).
 * Outputs (any subset): dist_out[nvec][n_refs], argmin_out[nvec], min_out[nvec]; all float64 / int32.
 */
int soap_classify(const void *X, const void *mult, const double *refw, const double *ref_norm2, int64_t nvec,
                  int dim, int n_refs, int kind, int power, int normalize_x, int dtype, double *dist_out,
                  int32_t *argmin_out, double *min_out, void *stream);

/* ---- transition counts of a classification: transitionMatrixFromSOAPClassification,
 * transitions/__init__.py:9-48.  states[n_frames][n_atoms] int32 (class index per frame and atom);
 * counts[n_classes][n_classes] uint64 is zeroed and then
 *     counts[states[f-window][a]][states[f][a]] += 1   for f in range(window, n_frames, stride), all a.
 * Entries outside [0, n_classes) are skipped.  Bit-exact.
 */
int soap_transition_counts(const int32_t *states, int64_t n_frames, int64_t n_atoms, int n_classes, int64_t window,
                           int64_t stride, uint64_t *counts, void *stream);

/* ---- residence times of a classification: calculateResidenceTimesFromClassification,
 * transitions/__init__.py:99-167.  For every atom and every initial frame in range(0, window, stride)
 * the series states[initial + k*window][atom] is walked; each change of state emits (old state, time
 * spent), the first interval of a series negated, and every series closes with -(time + window).
 * Events are appended in no particular order (the reference sorts them per state) to
 * ev_state[] / ev_time[]; *n_events receives their number.  capacity must be at least
 * soap_residence_capacity().  Integer work, bit-exact.
 */
int64_t soap_residence_capacity(int64_t n_frames, int64_t n_atoms, int64_t window, int64_t stride);
int soap_residence_events(const int32_t *states, int64_t n_frames, int64_t n_atoms, int n_classes, int64_t window,
                          int64_t stride, int32_t *ev_state, int64_t *ev_time, int64_t capacity, uint64_t *n_events,
                          void *stream);

/* ---- strided row copy between host and device on a stream (cudaMemcpy2DAsync): used to upload atom
 * tiles x[:, lo:hi, :] of a pinned host trajectory while the previous tile is being processed.
 */
int soap_copy_rows(void *dst, size_t dst_pitch, const void *src, size_t src_pitch, size_t row_bytes, size_t rows,
                   int to_device /* 1: host -> device, 0: device -> host, 2: device -> device (peer mapped included) */,
                   void *stream);

/* ---- SMs left free by the persistent kernels (soap_timelag runs one CTA per SM that owns the whole shared memory:
 * a collective launched on another stream cannot start beside it).  soap_reserve_sms(n) makes them use n SMs fewer,
 * so that e.g. the NCCL all-gather of the previous step's results overlaps the next sweep; 0 restores the default.
 */
int soap_reserve_sms(int n);

/* ---- result gather: NCCL all-gathers the ranks' t[F-w][A/G] slabs into
 * buf[G][rows][width]; this turns them into the out[rows][G * width] array of the reference layout
 * (timedSOAP[frame][atom], src/SOAPify/analysis.py:53) in one pass.
 */
int soap_interleave_shards(const double *buf, double *out, int64_t n_shards, int64_t rows, int64_t width, void *stream);

/* ---- synthetic data for the benchmarks (counter-based hash RNG, reproducible) ------------
 * fills n elements of `dtype` with uniform [0,1) values derived from (seed, element index).
 */
int soap_fill_uniform(void *out, int64_t n, uint64_t seed, int dtype, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SOAP_B200_H */
