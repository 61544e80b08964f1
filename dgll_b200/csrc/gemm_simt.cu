// gemm_simt.cu — exact-fp32 dense transform C = epi(op(A).op(B) + bias) (+C).
//
// This is the PARITY path of the dense per-layer transform X.W
// (torch.mm at dgll/nn/Convolution/gcnconv.py:30, Evaluation/PPI/gcn_model.py:70,
// gatconv.py:117; recomputed per edge in gcn_fused_kernel.cu:46-54): plain fp32
// FMAs so results stay within 1e-5 of the reference.  The throughput path is the
// tcgen05 kernel in gemm_tcgen05.cu (precision=1).
// 64x64 output tile, BK=16, 256 threads, 4x4 register micro-tile, smem staged.
#include "common.cuh"

namespace dgllb {

constexpr int BM = 64, BN = 64, BK = 16;

template <bool TA, bool TB>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const float* __restrict__ A, long long lda, const float* __restrict__ B, long long ldb,
                 float* __restrict__ C, long long ldc, long long M, long long N, long long K,
                 const float* __restrict__ bias, int epi, int accumulate) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long m0 = static_cast<long long>(blockIdx.y) * BM;
    const long long n0 = static_cast<long long>(blockIdx.x) * BN;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (long long k0 = 0; k0 < K; k0 += BK) {
        // stage A tile (BM x BK) and B tile (BK x BN): 1024 elements each, 4 per thread
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int idx = tid + t * 256;
            {
                // A: choose the fast index along the contiguous dimension
                int mm, kk;
                if (TA) { mm = idx & (BM - 1); kk = idx >> 6; }   // A stored [K, M]
                else    { kk = idx & (BK - 1); mm = idx >> 4; }   // A stored [M, K]
                const long long gm = m0 + mm, gk = k0 + kk;
                float v = 0.f;
                if (gm < M && gk < K) v = TA ? __ldg(A + gk * lda + gm) : __ldg(A + gm * lda + gk);
                As[kk][mm] = v;
            }
            {
                int nn, kk;
                if (TB) { kk = idx & (BK - 1); nn = idx >> 4; }   // B stored [N, K]
                else    { nn = idx & (BN - 1); kk = idx >> 6; }   // B stored [K, N]
                const long long gn = n0 + nn, gk = k0 + kk;
                float v = 0.f;
                if (gn < N && gk < K) v = TB ? __ldg(B + gn * ldb + gk) : __ldg(B + gk * ldb + gn);
                Bs[kk][nn] = v;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const long long gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            float v = acc[i][j];
            if (bias) v += __ldg(bias + gn);
            if (accumulate) v += C[gm * ldc + gn];
            if (epi & DGLLB_EPI_RELU) v = fmaxf(v, 0.f);
            if (epi & DGLLB_EPI_ELU) v = v > 0.f ? v : expm1f(v);
            C[gm * ldc + gn] = v;
        }
    }
}

int gemm_simt(const float* A, long long lda, int transA, const float* B, long long ldb, int transB, float* C,
              long long ldc, long long M, long long N, long long K, const float* bias, int epi,
              int accumulate, cudaStream_t st) {
    if (M == 0 || N == 0) return DGLLB_OK;
    dim3 grid(static_cast<unsigned>((N + BN - 1) / BN), static_cast<unsigned>((M + BM - 1) / BM));
    DGLLB_REQUIRE(grid.y < 65536u * 1024u, "gemm: M too large");
    if (grid.y > 65535u) {
        // split along M so gridDim.y stays legal
        const long long rows_per = 65535ll * BM;
        for (long long m = 0; m < M; m += rows_per) {
            const long long mm = M - m < rows_per ? M - m : rows_per;
            const float* Ap = transA ? A + m : A + m * lda;
            int rc = gemm_simt(Ap, lda, transA, B, ldb, transB, C + m * ldc, ldc, mm, N, K, bias, epi,
                               accumulate, st);
            if (rc != DGLLB_OK) return rc;
        }
        return DGLLB_OK;
    }
    if (transA && transB)
        gemm_simt_kernel<true, true><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, bias, epi, accumulate);
    else if (transA)
        gemm_simt_kernel<true, false><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, bias, epi, accumulate);
    else if (transB)
        gemm_simt_kernel<false, true><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, bias, epi, accumulate);
    else
        gemm_simt_kernel<false, false><<<grid, 256, 0, st>>>(A, lda, B, ldb, C, ldc, M, N, K, bias, epi, accumulate);
    DGLLB_LAUNCH_CHECK();
    return DGLLB_OK;
}

}  // namespace dgllb
