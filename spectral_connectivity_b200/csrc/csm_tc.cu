// Tensor-core (tcgen05) cross-spectral matrix GEMM -- placeholder until the UMMA kernel lands.
#include "sc_common.cuh"

int sc_csm_tc_supported(int64_t R, int64_t S) {
    (void)R; (void)S;
    return 0;
}

int sc_csm_tc_launch(const float* xp, int64_t B, int64_t F, int64_t R, int64_t S, float scale, void* out,
                     cudaStream_t st) {
    (void)xp; (void)B; (void)F; (void)R; (void)S; (void)scale; (void)out; (void)st;
    sc_set_error("sc_csm: tensor-core path not built");
    return SC_ERR_UNSUPPORTED;
}
