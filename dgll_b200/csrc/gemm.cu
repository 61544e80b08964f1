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
    DGLLB_REQUIRE(precision >= 0 && precision <= 3, "gemm: unknown precision %d", precision);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (precision == 3)
        return gemm_tf32x3(A, lda, transA, B, ldb, transB, C, ldc, M, N, K, bias, epilogue, accumulate, st);
    if (precision == 2)
        return gemm_tf32(A, lda, transA, B, ldb, transB, C, ldc, M, N, K, bias, epilogue, accumulate, st);
    if (precision == 1)
        return gemm_tcgen05(A, lda, transA, B, ldb, transB, C, ldc, M, N, K, bias, epilogue, accumulate, st);
    // precision 0 = fp32-grade results (north_star's 1e-5 bar).  Products large enough to amortise a persistent launch
    // run as 3xTF32 on the tensor cores (measured 1-4e-6 of max|ref| against 0.6-1.1e-6 for the FMA kernel, 3-6x faster,
    // 2x the library's exact-fp32 SGEMM; profiles/r02_gemm_tf32x3.jsonl); small ones and option gemm_kernel=5 take the
    // exact SIMT FMA kernel.
    // (its error grows with the length of one accumulation chain — the tensor core's fp32 add truncates — so a long
    // reduction qualifies only when the output is small enough for the kernel's split-K to cut the chains: K <= 2,048 or
    // at most 74 output tiles)
    if (opt_get(OPT_GEMM_KERNEL) != 5 && static_cast<double>(M) * static_cast<double>(N) * static_cast<double>(K) >= 2.5e7 &&
        K >= 32 && N >= 16 && (K <= 2048 || ((M + 127) / 128) * ((N + 127) / 128) <= 74))
        return gemm_tf32x3(A, lda, transA, B, ldb, transB, C, ldc, M, N, K, bias, epilogue, accumulate, st);
    return gemm_simt(A, lda, transA, B, ldb, transB, C, ldc, M, N, K, bias, epilogue, accumulate, st);
}
