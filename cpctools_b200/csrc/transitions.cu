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

// A thread owns an atom and walks its frame steps.  Slow dynamics means long runs of the SAME transition (mostly
// i -> i): the thread counts the run in a register and touches the histogram only when the transition changes -- no
// warp vote, no atomic per step (the first version elected one lane per distinct cell with __match_any_sync every
// step: 0.22 of the HBM peak on the bytes actually moved).  With window == stride the `from` state of a step is the
// `to` state of the previous one and stays in a register: every state is read once.
template <bool CHAIN>  // CHAIN: window == stride
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
  // thread = atom (coalesced rows), blockIdx.y = a slice of the frame steps window, window+stride, ...
  const long long n_steps = (n_frames - window + stride - 1) / stride;
  const long long per_y = (n_steps + gridDim.y - 1) / gridDim.y;
  const long long s0 = (long long)blockIdx.y * per_y, s1 = min(n_steps, s0 + per_y);
  for (long long atom = (long long)blockIdx.x * kTrThreads + threadIdx.x; atom < n_atoms && s0 < s1;
       atom += (long long)gridDim.x * kTrThreads) {
    int last_key = -1;
    unsigned run = 0;
    auto flush = [&]() {
      if (run && last_key >= 0) {
        if (use_smem)
          atomicAdd(&hist[last_key], run);
        else
          atomicAdd(&counts[last_key], (unsigned long long)run);
      }
    };
    auto count = [&](int from, int to) {
      const bool ok = (unsigned)from < (unsigned)n_classes && (unsigned)to < (unsigned)n_classes;
      const int key = ok ? from * n_classes + to : -1;
      if (key != last_key) {
        flush();
        last_key = key;
        run = 0;
      }
      ++run;
    };
    constexpr int kBatch = 8;  // independent loads in flight per thread
    const int *col = states + atom;
    int prev = CHAIN ? col[(window + s0 * stride - window) * n_atoms] : 0;
    long long s = s0;
    for (; s + kBatch <= s1; s += kBatch) {
      int from[kBatch], to[kBatch];
#pragma unroll
      for (int j = 0; j < kBatch; ++j) {
        const long long frame = window + (s + j) * stride;
        to[j] = col[frame * n_atoms];
        if (!CHAIN) from[j] = col[(frame - window) * n_atoms];
      }
#pragma unroll
      for (int j = 0; j < kBatch; ++j) {
        count(CHAIN ? prev : from[j], to[j]);
        prev = to[j];
      }
    }
    for (; s < s1; ++s) {
      const long long frame = window + s * stride;
      const int to = col[frame * n_atoms];
      count(CHAIN ? prev : col[(frame - window) * n_atoms], to);
      prev = to;
    }
    flush();
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
  auto step = [&](int now) {
    time += window;
    if (now != state) {
      emit(state, first ? -time : time);
      first = false;
      state = now;
      time = 0;
    }
  };
  // the walk is a dependent chain only through `state`: 8 frames are fetched at once so that a thread keeps 8
  // loads in flight (one load per step leaves the memory system ~6 % busy; 16 measured slower, 0.42 against 0.30 ms
  // for 2000 x 100 000 states).  emit() adds the CONSTANT 1 to the cursor, which the compiler turns into one atomic
  // per warp (vote + elect); reserving the slots of a whole batch with one variable-size atomic per thread defeats
  // that and measured 20x slower (8.7 M same-address atomics).  With one series per atom the kernel is limited by
  // the number of series (100 000 threads keep ~20 kB per SM in flight), not by the arithmetic.
  constexpr int kBatch = 8;
  long long f = window + initial;
  for (; f + (kBatch - 1) * window < n_frames; f += kBatch * window) {
    int nxt[kBatch];
#pragma unroll
    for (int j = 0; j < kBatch; ++j) nxt[j] = states[(f + j * window) * n_atoms + atom];
#pragma unroll
    for (int j = 0; j < kBatch; ++j) step(nxt[j]);
  }
  for (; f < n_frames; f += window) step(states[f * n_atoms + atom]);
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
  const int64_t n_steps = (n_frames - window + stride - 1) / stride;
  int64_t gx = (n_atoms + kTrThreads - 1) / kTrThreads;
  const int64_t cap = (int64_t)num_sms() * 8;
  gx = gx > cap ? cap : gx;
  int64_t gy = (cap + gx - 1) / gx;  // enough CTAs to fill the device; every CTA flushes one histogram
  gy = gy > n_steps ? n_steps : gy;
  gy = gy < 1 ? 1 : (gy > 65535 ? 65535 : gy);
  if (window == stride)
    transition_counts_kernel<true><<<dim3((unsigned)gx, (unsigned)gy), kTrThreads, smem, s>>>(
        states, n_frames, n_atoms, n_classes, window, stride, reinterpret_cast<unsigned long long *>(counts));
  else
    transition_counts_kernel<false><<<dim3((unsigned)gx, (unsigned)gy), kTrThreads, smem, s>>>(
        states, n_frames, n_atoms, n_classes, window, stride, reinterpret_cast<unsigned long long *>(counts));
  return check_launch("transition_counts_kernel");
}
