// bst_tc.cu — BST transformer block with the projections, the FFN and the weight gradients on
// tcgen05 tensor cores (placeholder until the kernels land: the entry points report "not built").
#include "bst.cuh"

namespace rk {

int bst_tc_bwd_ctas(int64_t, int) { return 1; }
int bst_tc_fwd(const BstParams&, int, float*, float*, int, int32_t*, cudaStream_t) {
    RK_CHECK_ARG(false, "bst: the tensor-core block is not built yet");
    return -1;
}
int bst_tc_bwd(const BstParams&, int, const float*, const float*, int, float*, float*, float*, int, int32_t*,
               cudaStream_t) {
    RK_CHECK_ARG(false, "bst: the tensor-core block is not built yet");
    return -1;
}

}  // namespace rk
