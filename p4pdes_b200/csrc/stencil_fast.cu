// stencil_fast.cu -- placeholder until the TMA-staged plane-marching kernel lands.
#include "kernels.h"
namespace p4b {
bool stencil_fast_eligible(const LevelDesc &) { return false; }
int launch_stencil_fast(cudaStream_t, const LevelDesc &, const StencilOp &, const Reducer &) {
    return fail(62, "fast stencil path not built");
}
}  // namespace p4b
