// placeholder, replaced below
#include "common.cuh"
namespace soapb {
size_t pairs_dmma_scratch(int64_t, int64_t, int) { return 16; }
int pairs_dmma(const double *, const double *, double *, int64_t, int64_t, int, int64_t, int, int, void *, cudaStream_t) {
  return set_error(SOAP_EUNSUPPORTED, "pairs_dmma: not built yet");
}
}
