// Backward instantiations: SHARP = true, HAS_MASK = false (see t2o_chain_kernels.cuh).
#include "t2o_chain_kernels.cuh"

namespace t2o {
int launch_bwd_sharp_nomask(int vec, bool small_chain, const BwdArgs &a, size_t smem, cudaStream_t stream) {
    return launch_bwd_sel<true, false>(vec, small_chain, a, smem, stream);
}
}  // namespace t2o
