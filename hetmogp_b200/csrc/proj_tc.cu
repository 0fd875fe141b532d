// Tensor-core (tcgen05) projection path -- placeholder until the kernels land; HMOGP_PREC_TC is rejected at create.
#include "common.cuh"
int hm_tc_available() { return 0; }
int hm_tc_prepare(cudaStream_t, const double*, const double*, void*, int, int, int) { return 0; }
int hm_tc_proj_fwd(cudaStream_t, const HmTasks&, const HmProjArgs&, const void*) {
    hm_set_error("tensor-core path not built");
    return HMOGP_ERR_ARG;
}
