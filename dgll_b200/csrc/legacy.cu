// legacy.cu — the reference's fused-GCN C ABI, re-implemented on the new kernels.
//
// Replaces launch_gcn_fused_kernel / launch_gcn_fused_kernel_backward_optimized
// (dgll/FusedKernel/gcn_fused_kernel.cu:190-287) and adds stream-ordered v2s.
// Forward  : S = X[:, :actual_F] . W[:actual_F, :]   (dense transform, once per node —
//            the reference recomputes it per edge per output column, .cu:46-54)
//            H = relu(A_hat . S)                     (CSR aggregation, relu fused)
// Backward : Gm = G * (H > 0);  AX = A_hat . X;  grad_W += AX^T . Gm
//            T = Gm . W^T;      grad_X += A_hat^T . T
//            (the true gradients; the reference kernel's are wrong)
// Workspace is taken from the stream-ordered allocator.
#include "common.cuh"
#include "internal.cuh"

namespace dgllb {

// Gm[r, c] = H[r, c] > 0 ? G[r, c] : 0     (G, H compact [N, Hd]; Gm has ld = ldm)
__global__ void relu_mask_kernel(const float* __restrict__ g, const float* __restrict__ h,
                                 float* __restrict__ out, long long ldm, long long n_rows, int Hd) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_rows * Hd) return;
    const long long r = i / Hd;
    const int c = static_cast<int>(i - r * Hd);
    out[r * ldm + c] = h[i] > 0.f ? g[i] : 0.f;
}

// vals_eff[e] = values[e] if e is within the first num_neighbors[row] edges of its row, else 0
__global__ void clamp_values_kernel(const int* __restrict__ row_ptr, const int* __restrict__ cnt,
                                    const float* __restrict__ values, float* __restrict__ out, int n_rows) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n_rows) return;
    const int b = row_ptr[row], e = row_ptr[row + 1];
    const int lim = b + max(0, cnt[row]);
    for (int i = b + lane; i < e; i += 32) out[i] = i < lim ? values[i] : 0.f;
}

static SpmmParams make_params(const int* row_ptr, const int* col, const float* vals, const float* X,
                              long long ldx, float* out, long long ldo, long long n, int F, int epi,
                              const int* row_cnt) {
    SpmmParams p;
    p.row_ptr = row_ptr; p.rp64 = 0; p.col = col; p.vals = vals; p.X = X; p.ldx = ldx;
    p.out = out; p.ldo = ldo; p.n_dst = n; p.F = F; p.n_slabs = 1; p.mean = 0;
    p.row_scale = nullptr; p.addend = nullptr; p.ld_add = 0; p.bias = nullptr; p.epi = epi;
    p.argmax = nullptr; p.heavy_items = nullptr; p.n_heavy_items = 0; p.chunk_edges = 0; p.heavy_ws = nullptr; p.ld_hws = 0;
    p.row_cnt = row_cnt;
    p.nnz_hint = -1;
    p.out_vec = 0;
    return p;
}

static inline size_t al256(size_t b) { return (b + 255) & ~static_cast<size_t>(255); }

}  // namespace dgllb

using namespace dgllb;

extern "C" int dgllb_gcn_fused_forward(const int* row_ptr, const int* col_idx, const float* values,
                                       const float* X, const float* W, float* H,
                                       const int* num_neighbors, int N, int F_padded, int actual_F,
                                       int H_dim, int total_nnz, void* stream) {
    DGLLB_REQUIRE(N >= 0 && H_dim >= 0 && actual_F >= 0 && actual_F <= F_padded && total_nnz >= 0,
                  "gcn_fused_forward: bad sizes N=%d F_padded=%d actual_F=%d H=%d nnz=%d", N, F_padded,
                  actual_F, H_dim, total_nnz);
    if (N == 0 || H_dim == 0) return DGLLB_OK;
    DGLLB_REQUIRE(row_ptr && X && W && H && (total_nnz == 0 || (col_idx && values)),
                  "gcn_fused_forward: null pointer");
    { DevInfo di_; int rc_ = get_devinfo(&di_); if (rc_ != DGLLB_OK) return rc_; }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* S = nullptr;
    const long long ldS = (H_dim + 3) & ~3;  // keep rows 16-byte aligned for the vector path
    DGLLB_CUDA_TRY(cudaMallocAsync(&S, sizeof(float) * static_cast<size_t>(N) * ldS, st));
    int rc = gemm_simt(X, F_padded, 0, W, H_dim, 0, S, ldS, N, H_dim, actual_F, nullptr, 0, 0, st);
    if (rc == DGLLB_OK) {
        SpmmParams p = make_params(row_ptr, col_idx, values, S, ldS, H, H_dim, N, H_dim, DGLLB_EPI_RELU,
                                   num_neighbors);
        rc = spmm_run(p, DGLLB_F32, false, nullptr, st);
    }
    cudaFreeAsync(S, st);
    return rc;
}

extern "C" int dgllb_gcn_fused_backward(const int* row_ptr, const int* col_idx, const float* values,
                                        const float* X, const float* W, const float* H,
                                        const float* grad_output, float* grad_W, float* grad_X,
                                        const int* num_neighbors, int N, int F_padded, int actual_F,
                                        int H_dim, int total_nnz, void* stream) {
    DGLLB_REQUIRE(N >= 0 && H_dim >= 0 && actual_F >= 0 && actual_F <= F_padded && total_nnz >= 0,
                  "gcn_fused_backward: bad sizes");
    if (N == 0 || H_dim == 0 || actual_F == 0) return DGLLB_OK;
    DGLLB_REQUIRE(row_ptr && X && W && H && grad_output && grad_W && grad_X &&
                      (total_nnz == 0 || (col_idx && values)),
                  "gcn_fused_backward: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long ldF = (actual_F + 3) & ~3;
    const long long ldH = (H_dim + 3) & ~3;
    const size_t nnz = static_cast<size_t>(total_nnz);
    const size_t bGm = al256(sizeof(float) * static_cast<size_t>(N) * ldH);
    const size_t bNF = al256(sizeof(float) * static_cast<size_t>(N) * ldF);
    const size_t bRp = al256(sizeof(int) * (static_cast<size_t>(N) + 1));
    const size_t bE = al256(sizeof(int) * (nnz + 1));
    char* ws = nullptr;
    DGLLB_CUDA_TRY(cudaMallocAsync(&ws, bGm + 2 * bNF + bRp + 3 * bE, st));
    float* Gm = reinterpret_cast<float*>(ws);
    float* AX = reinterpret_cast<float*>(ws + bGm);
    float* T = reinterpret_cast<float*>(ws + bGm + bNF);
    int* t_row_ptr = reinterpret_cast<int*>(ws + bGm + 2 * bNF);
    int* t_col = reinterpret_cast<int*>(ws + bGm + 2 * bNF + bRp);
    float* t_val = reinterpret_cast<float*>(ws + bGm + 2 * bNF + bRp + bE);
    float* v_eff = reinterpret_cast<float*>(ws + bGm + 2 * bNF + bRp + 2 * bE);

    int rc = DGLLB_OK;
    do {
        const long long nel = static_cast<long long>(N) * H_dim;
        relu_mask_kernel<<<static_cast<unsigned>((nel + 255) / 256), 256, 0, st>>>(grad_output, H, Gm, ldH, N,
                                                                                   H_dim);
        g_launch_count.fetch_add(1);
        // AX = A_hat . X[:, :actual_F]
        SpmmParams p1 = make_params(row_ptr, col_idx, values, X, F_padded, AX, ldF, N, actual_F, 0,
                                    num_neighbors);
        if ((rc = spmm_run(p1, DGLLB_F32, false, nullptr, st)) != DGLLB_OK) break;
        // grad_W[:actual_F, :] += AX^T . Gm
        if ((rc = gemm_simt(AX, ldF, 1, Gm, ldH, 0, grad_W, H_dim, actual_F, H_dim, N, nullptr, 0, 1, st)) !=
            DGLLB_OK)
            break;
        // T = Gm . W[:actual_F, :]^T
        if ((rc = gemm_simt(Gm, ldH, 0, W, H_dim, 1, T, ldF, N, actual_F, H_dim, nullptr, 0, 0, st)) != DGLLB_OK)
            break;
        // grad_X[:, :actual_F] += A_hat^T . T
        const float* vals_for_t = values;
        if (num_neighbors && total_nnz > 0) {
            clamp_values_kernel<<<static_cast<unsigned>((static_cast<long long>(N) * 32 + 255) / 256), 256, 0,
                                  st>>>(row_ptr, num_neighbors, values, v_eff, N);
            g_launch_count.fetch_add(1);
            vals_for_t = v_eff;
        }
        if ((rc = dgllb_csr_transpose(row_ptr, 0, col_idx, vals_for_t, N, N, total_nnz, t_row_ptr, t_col, t_val,
                                      nullptr, stream)) != DGLLB_OK)
            break;
        SpmmParams p2 = make_params(t_row_ptr, t_col, t_val, T, ldF, grad_X, F_padded, N, actual_F, 0, nullptr);
        p2.addend = grad_X;  // accumulate into the caller-zeroed buffer (same element, same thread)
        p2.ld_add = F_padded;
        if ((rc = spmm_run(p2, DGLLB_F32, false, nullptr, st)) != DGLLB_OK) break;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            set_error("gcn_fused_backward: %s", cudaGetErrorString(e));
            rc = DGLLB_ERR_CUDA;
        }
    } while (0);
    cudaFreeAsync(ws, st);
    return rc;
}

// ------------------------------------------------ reference-compatible symbols --
extern "C" void launch_gcn_fused_kernel(const int* row_ptr, const int* col_idx, const float* values,
                                        const float* X, const float* W, float* H,
                                        const int* num_neighbors, int N, int F_padded, int actual_F,
                                        int H_dim, int total_nnz) {
    int rc = dgllb_gcn_fused_forward(row_ptr, col_idx, values, X, W, H, num_neighbors, N, F_padded, actual_F,
                                     H_dim, total_nnz, nullptr);
    if (rc == DGLLB_OK) {
        cudaError_t e = cudaDeviceSynchronize();  // the reference blocks here too (.cu:229)
        if (e != cudaSuccess) {
            set_error("launch_gcn_fused_kernel: %s", cudaGetErrorString(e));
            rc = DGLLB_ERR_CUDA;
        }
    }
    if (rc != DGLLB_OK) fprintf(stderr, "dgll_b200: launch_gcn_fused_kernel failed: %s\n", dgllb_last_error());
}

extern "C" void launch_gcn_fused_kernel_backward_optimized(
    const int* row_ptr, const int* col_idx, const float* values, const float* X, const float* W,
    const float* grad_output, float* grad_W, float* grad_X, const int* num_neighbors, int N, int F_padded,
    int actual_F, int H_dim, int total_nnz) {
    // the legacy signature does not carry H: recompute the forward for the ReLU mask
    float* Htmp = nullptr;
    int rc = DGLLB_OK;
    if (N > 0 && H_dim > 0) {
        cudaError_t e = cudaMallocAsync(&Htmp, sizeof(float) * static_cast<size_t>(N) * H_dim, nullptr);
        if (e != cudaSuccess) {
            set_error("launch_gcn_fused_kernel_backward_optimized: %s", cudaGetErrorString(e));
            rc = DGLLB_ERR_CUDA;
        }
    }
    if (rc == DGLLB_OK)
        rc = dgllb_gcn_fused_forward(row_ptr, col_idx, values, X, W, Htmp, num_neighbors, N, F_padded, actual_F,
                                     H_dim, total_nnz, nullptr);
    if (rc == DGLLB_OK)
        rc = dgllb_gcn_fused_backward(row_ptr, col_idx, values, X, W, Htmp, grad_output, grad_W, grad_X,
                                      num_neighbors, N, F_padded, actual_F, H_dim, total_nnz, nullptr);
    if (Htmp) cudaFreeAsync(Htmp, nullptr);
    if (rc == DGLLB_OK) {
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            set_error("launch_gcn_fused_kernel_backward_optimized: %s", cudaGetErrorString(e));
            rc = DGLLB_ERR_CUDA;
        }
    }
    if (rc != DGLLB_OK)
        fprintf(stderr, "dgll_b200: launch_gcn_fused_kernel_backward_optimized failed: %s\n", dgllb_last_error());
}
