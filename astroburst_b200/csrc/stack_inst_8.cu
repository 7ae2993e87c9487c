// Explicit instantiations of the exact-frame-count stack kernels (one group per file so
// the build compiles them in parallel).  Generated list, see stack.cu for the dispatcher.
#include "stack_kernels.cuh"

namespace abk {
template ab_status launch_tile<40>(ab_ctx *, const FrameTable &, size_t, float, float, unsigned,
                                      float *, unsigned long long *, cudaStream_t);
template ab_status launch_tile<36>(ab_ctx *, const FrameTable &, size_t, float, float, unsigned,
                                      float *, unsigned long long *, cudaStream_t);
}  // namespace abk
