#include "fit_dmma_launch.h"
namespace fastcheb {
bool fit_dmma_applicable(int, int, long long) { return false; }
int fit_dmma_1d(const DeviceInfo*, int, int, long long, int, const void*, void*, int, cudaStream_t) {
  return fail(FASTCHEB_EUNSUPPORTED, "fit_dmma not built");
}
}  // namespace fastcheb
