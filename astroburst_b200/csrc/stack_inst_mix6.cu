// Explicit instantiation of the P = 32 stack kernels with 3/4 of the compare-exchanges in the
// IMAD form (AB_STACK_MIX=6; A/B measurement only, see stack_kernels.cuh).
#ifdef AB_DEV_VARIANTS  // developer A/B builds only
#include "stack_kernels.cuh"

namespace abk {
template ab_status launch_mix<32, 3, 4>(ab_ctx *, const FrameTable &, size_t, float, float, unsigned,
                                        float *, unsigned long long *, cudaStream_t);
}  // namespace abk
#endif
