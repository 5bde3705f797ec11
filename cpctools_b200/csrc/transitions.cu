// transitions.cu -- f4: unnormalised transition matrix and residence times of a classification.
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

// Residence times of a classification: calculateResidenceTimesFromClassification
// (src/SOAPify/transitions/__init__.py:99-167).  One thread owns one series = (atom, initial frame):
// it walks the frames initial + k*window and emits (state, time) whenever the state changes, the
// first interval of a series negated, plus the closing  -(time + window)  of the open interval.
// Consecutive threads are consecutive atoms, so every frame row is read coalesced.  Events go to one
// global list (slot = atomic cursor); the reference returns them SORTED per state, so their order of
// production is irrelevant (the host sorts on the device).
__global__ void __launch_bounds__(256)
    residence_events_kernel(const int *__restrict__ states, long long n_frames, long long n_atoms, int n_classes,
                            long long window, long long stride, int n_init, int *__restrict__ ev_state,
                            long long *__restrict__ ev_time, long long capacity, unsigned long long *__restrict__ cursor) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_atoms * n_init) return;
  const long long atom = i % n_atoms;
  const long long initial = (i / n_atoms) * stride;
  auto emit = [&](int st, long long t) {
    const unsigned long long slot = atomicAdd(cursor, 1ull);
    if ((long long)slot < capacity) {
      ev_state[slot] = st;
      ev_time[slot] = t;
    }
  };
  long long time = 0;
  int state = states[initial * n_atoms + atom];
  bool first = true;
  for (long long f = window + initial; f < n_frames; f += window) {
    time += window;
    const int now = states[f * n_atoms + atom];
    if (now != state) {
      emit(state, first ? -time : time);
      first = false;
      state = now;
      time = 0;
    }
  }
  emit(state, -time - window);  // the open interval at the end of the series
}

}  // namespace soapb

using namespace soapb;

extern "C" int64_t soap_residence_capacity(int64_t n_frames, int64_t n_atoms, int64_t window, int64_t stride) {
  if (n_frames <= 0 || n_atoms <= 0 || window < 1 || stride < 1) return 0;
  const int64_t n_init = (window + stride - 1) / stride;            // len(range(0, window, stride))
  return n_atoms * n_init * (n_frames / window + 2);                 // visited frames + the closing event
}

extern "C" int soap_residence_events(const int32_t *states, int64_t n_frames, int64_t n_atoms, int n_classes,
                                     int64_t window, int64_t stride, int32_t *ev_state, int64_t *ev_time,
                                     int64_t capacity, uint64_t *n_events, void *stream) {
  SOAP_REQUIRE(n_frames > 0 && n_atoms > 0 && n_classes > 0, "soap_residence_events: bad shape");
  SOAP_REQUIRE(window >= 1 && stride >= 1 && stride <= window && window <= n_frames,
               "soap_residence_events: bad window / stride");
  SOAP_REQUIRE(states && ev_state && ev_time && n_events, "soap_residence_events: null buffer");
  SOAP_REQUIRE(capacity >= soap_residence_capacity(n_frames, n_atoms, window, stride),
               "soap_residence_events: capacity %lld is below soap_residence_capacity()", (long long)capacity);
  auto s = static_cast<cudaStream_t>(stream);
  SOAP_CUDA(cudaMemsetAsync(n_events, 0, sizeof(uint64_t), s));
  const int64_t n_init = (window + stride - 1) / stride;
  const int64_t series = n_atoms * n_init;
  SOAP_REQUIRE(n_init < (1ll << 31) && (series + 255) / 256 < (1ll << 31), "soap_residence_events: too many series");
  residence_events_kernel<<<(unsigned)((series + 255) / 256), 256, 0, s>>>(
      states, n_frames, n_atoms, n_classes, window, stride, (int)n_init, ev_state, reinterpret_cast<long long *>(ev_time),
      capacity, reinterpret_cast<unsigned long long *>(n_events));
  return check_launch("residence_events_kernel");
}

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
