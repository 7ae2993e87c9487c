// Explicit instantiations of the exact-frame-count stack kernels (one group per file so
// the build compiles them in parallel).  Generated list, see stack.cu for the dispatcher.
#include "stack_kernels.cuh"

namespace abk {
template ab_status launch_tile<8>(ab_ctx *, const FrameTable &, size_t, float, float, unsigned,
                                      float *, unsigned long long *, cudaStream_t);
template ab_status launch_tile<7>(ab_ctx *, const FrameTable &, size_t, float, float, unsigned,
                                      float *, unsigned long long *, cudaStream_t);
template ab_status launch_tile<6>(ab_ctx *, const FrameTable &, size_t, float, float, unsigned,
                                      float *, unsigned long long *, cudaStream_t);
template ab_status launch_tile<5>(ab_ctx *, const FrameTable &, size_t, float, float, unsigned,
                                      float *, unsigned long long *, cudaStream_t);
template ab_status launch_tile<4>(ab_ctx *, const FrameTable &, size_t, float, float, unsigned,
                                      float *, unsigned long long *, cudaStream_t);
template ab_status launch_tile<3>(ab_ctx *, const FrameTable &, size_t, float, float, unsigned,
                                      float *, unsigned long long *, cudaStream_t);
template ab_status launch_tile<2>(ab_ctx *, const FrameTable &, size_t, float, float, unsigned,
                                      float *, unsigned long long *, cudaStream_t);
template ab_status launch_tile<1>(ab_ctx *, const FrameTable &, size_t, float, float, unsigned,
                                      float *, unsigned long long *, cudaStream_t);
}  // namespace abk
