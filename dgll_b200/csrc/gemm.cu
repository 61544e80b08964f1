// gemm.cu — public entry of the dense per-layer transform (dispatch only).
#include "common.cuh"
#include "internal.cuh"

using namespace dgllb;

extern "C" int dgllb_gemm_f32(const float* A, int64_t lda, int transA, const float* B, int64_t ldb,
                              int transB, float* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                              const float* bias, int epilogue, int accumulate, int precision,
                              void* stream) {
    DGLLB_REQUIRE(M >= 0 && N >= 0 && K >= 0, "gemm: negative size");
    if (M == 0 || N == 0) return DGLLB_OK;
    DGLLB_REQUIRE(C && (K == 0 || (A && B)), "gemm: null pointer");
    DGLLB_REQUIRE(ldc >= N, "gemm: ldc < N");
    DGLLB_REQUIRE(lda >= (transA ? M : K) && ldb >= (transB ? K : N), "gemm: leading dimension too small");
    DGLLB_REQUIRE(precision >= 0 && precision <= 2, "gemm: unknown precision %d", precision);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (precision == 2)
        return gemm_tf32(A, lda, transA, B, ldb, transB, C, ldc, M, N, K, bias, epilogue, accumulate, st);
    if (precision == 1)
        return gemm_tcgen05(A, lda, transA, B, ldb, transB, C, ldc, M, N, K, bias, epilogue, accumulate, st);
    return gemm_simt(A, lda, transA, B, ldb, transB, C, ldc, M, N, K, bias, epilogue, accumulate, st);
}
