// transitions.cu -- f4: unnormalised transition matrix of a classification.
//
// Reference semantics: transitionMatrixFromSOAPClassification
// (src/SOAPify/transitions/__init__.py:9-48):
//     for frame in range(window, nframes, stride): for atom: M[state[frame-window][atom]][state[frame][atom]] += 1
// Integer work, bit-exact: a per-CTA shared-memory histogram (class pairs), flushed with 64-bit
// global atomics.  Memory-bound on the int32 state array (each state is read twice, 4 B each).
#include "common.cuh"

namespace soapb {

constexpr int kTrThreads = 256;
constexpr int kTrMaxShared = 8192;  // class pairs kept in shared memory (90 x 90 classes)

__global__ void __launch_bounds__(kTrThreads)
    transition_counts_kernel(const int *__restrict__ states, long long n_frames, long long n_atoms, int n_classes,
                             long long window, long long stride, unsigned long long *__restrict__ counts) {
  extern __shared__ unsigned int hist[];
  const int pairs = n_classes * n_classes;
  const bool use_smem = pairs <= kTrMaxShared;
  if (use_smem) {
    for (int i = threadIdx.x; i < pairs; i += kTrThreads) hist[i] = 0u;
    __syncthreads();
  }
  const long long n_steps = (n_frames - window + stride - 1) / stride;  // frames window, window+stride, ...
  const long long total = n_steps * n_atoms;
  const long long grid_stride = (long long)gridDim.x * kTrThreads;
  for (long long i = (long long)blockIdx.x * kTrThreads + threadIdx.x; i < total; i += grid_stride) {
    const long long step = i / n_atoms, atom = i - step * n_atoms;
    const long long frame = window + step * stride;
    const int from = states[(frame - window) * n_atoms + atom], to = states[frame * n_atoms + atom];
    if ((unsigned)from < (unsigned)n_classes && (unsigned)to < (unsigned)n_classes) {
      if (use_smem)
        atomicAdd(&hist[from * n_classes + to], 1u);
      else
        atomicAdd(&counts[(size_t)from * n_classes + to], 1ull);
    }
  }
  if (use_smem) {
    __syncthreads();
    for (int i = threadIdx.x; i < pairs; i += kTrThreads)
      if (hist[i]) atomicAdd(&counts[i], (unsigned long long)hist[i]);
  }
}

}  // namespace soapb

using namespace soapb;

extern "C" int soap_transition_counts(const int32_t *states, int64_t n_frames, int64_t n_atoms, int n_classes,
                                      int64_t window, int64_t stride, uint64_t *counts, void *stream) {
  SOAP_REQUIRE(n_frames > 0 && n_atoms > 0 && n_classes > 0, "soap_transition_counts: bad shape");
  SOAP_REQUIRE(window >= 1 && stride >= 1 && window <= n_frames, "soap_transition_counts: bad window / stride");
  SOAP_REQUIRE(states && counts, "soap_transition_counts: null buffer");
  SOAP_REQUIRE((int64_t)n_classes * n_classes < (1ll << 31), "soap_transition_counts: too many classes");
  auto s = static_cast<cudaStream_t>(stream);
  SOAP_CUDA(cudaMemsetAsync(counts, 0, (size_t)n_classes * n_classes * sizeof(uint64_t), s));
  if (window >= n_frames) return SOAP_OK;  // range(window, nframes, stride) is empty
  const int pairs = n_classes * n_classes;
  const size_t smem = pairs <= kTrMaxShared ? (size_t)pairs * 4 : 0;
  const int64_t total = ((n_frames - window + stride - 1) / stride) * n_atoms;
  int64_t grid = (total + kTrThreads * 8 - 1) / (kTrThreads * 8);
  const int64_t cap = (int64_t)num_sms() * 8;
  grid = grid < 1 ? 1 : (grid > cap ? cap : grid);
  transition_counts_kernel<<<(unsigned)grid, kTrThreads, smem, s>>>(states, n_frames, n_atoms, n_classes, window, stride,
                                                                    reinterpret_cast<unsigned long long *>(counts));
  return check_launch("transition_counts_kernel");
}
