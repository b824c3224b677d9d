#include "lm.h"
namespace stl {
void lm_free(LmState &lm) { (void)lm; }
cudaError_t lm_associate(const DevPack &, const DevWork &, const DevParams &, LmState &, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t lm_linearize(const DevPack &, const DevParams &, LmState &, const double *, int, double *, cudaStream_t) { return cudaErrorNotSupported; }
}
