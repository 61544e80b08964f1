// gemm_tcgen05.cu — tensor-core dense transform (placeholder until the tcgen05 kernel lands).
#include "common.cuh"
#include "internal.cuh"

namespace dgllb {

int gemm_tcgen05(const float*, long long, int, const float*, long long, int, float*, long long, long long,
                 long long, long long, const float*, int, int, cudaStream_t) {
    set_error("gemm: precision=1 (tcgen05) is not built yet");
    return DGLLB_ERR_UNSUPPORTED;
}

}  // namespace dgllb
