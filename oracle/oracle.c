/*
 * oracle.c — CPU restatement of the reference's neighbourhood-aggregation path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under dgll_b200/ may import, link or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, as the checker and the timed CPU baseline.
 *
 * Every function cites the reference lines it restates (paths relative to the
 * reference tree).  Arithmetic is plain scalar C; sums accumulate in double so
 * the oracle is a tighter truth than either fp32 implementation (tolerance in
 * the tests: 1e-5 relative for fp32 paths, 1e-2 for bf16 paths; integer/byte
 * results must match bit for bit).
 *
 * Pinning: the reference has no numeric tests or golden vectors for this path
 * (SURVEY.md §4, §8c), so this file is pinned by (1) tests/golden/*.npz, produced
 * by running the reference's own Python modules in the build container
 * (oracle/gen_golden.py), and (2) on the GPU box, the reference's own forward
 * kernel compiled from its source into oracle/_ref (oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_SUM 0
#define ORC_MEAN 1
#define ORC_MAX 2
#define ORC_EPI_RELU 1
#define ORC_EPI_ELU 2

static float orc_epi(double v, int epi) {
    if (epi & ORC_EPI_RELU) v = v > 0.0 ? v : 0.0;
    if (epi & ORC_EPI_ELU) v = v > 0.0 ? v : expm1(v);
    return (float)v;
}

/*
 * out = epi(row_scale * reduce_e(values[e] * X[col[e]]) + addend + bias)
 * Restates: torch.spmm(adj, support) dgll/nn/Convolution/gcnconv.py:31;
 * torch.sparse.mm(adj_matrix, support) Evaluation/PPI/gcn_model.py:76;
 * the edge loop gcn_fused_kernel.cu:41-57; neighbor_feature.mean/sum/max(dim=1)
 * sageconv.py:32-38 (as intended); scatter(reduce=add/mean/max)
 * GlobalPooling/Pooling.py:37,59,81 (col == NULL: segment reduce);
 * DGL copy_u/sum|mean (restated from public semantics):
 * mean of an empty row = 0, max of an empty row = 0.
 */
void orc_spmm_csr(const int64_t* row_ptr, const int32_t* col, const float* values, const float* X,
                  int64_t ldx, float* out, int64_t ldo, int64_t n_dst, int F, int reduce,
                  const float* row_scale, const float* addend, int64_t ld_add, const float* bias,
                  int epi, int32_t* argmax) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n_dst; ++i) {
        const int64_t b = row_ptr[i], e = row_ptr[i + 1];
        const int64_t deg = e - b;
        for (int f = 0; f < F; ++f) {
            double acc = reduce == ORC_MAX ? -INFINITY : 0.0;
            int32_t am = -1;
            for (int64_t k = b; k < e; ++k) {
                const int64_t c = col ? (int64_t)col[k] : k;
                const double v = (double)(values ? values[k] : 1.0f) * (double)X[c * ldx + f];
                if (reduce == ORC_MAX) {
                    if (v > acc) { acc = v; am = (int32_t)k; }
                } else {
                    acc += v;
                }
            }
            if (reduce == ORC_MAX && deg == 0) acc = 0.0;
            if (reduce == ORC_MEAN) acc = deg > 0 ? acc / (double)deg : 0.0;
            if (row_scale) acc *= (double)row_scale[i];
            if (addend) acc += (double)addend[i * ld_add + f];
            if (bias) acc += (double)bias[f];
            out[i * ldo + f] = orc_epi(acc, epi);
            if (argmax) argmax[i * (int64_t)F + f] = am;
        }
    }
}

/*
 * H = relu(A_hat (X W)) exactly as the reference kernel walks it
 * (dgll/FusedKernel/gcn_fused_kernel.cu:26-70): for every row, output column j
 * and edge, z = sum_f X[col, f] * W[f, j] over f < actual_F (X has stride
 * F_padded), sum += a_val * z, output = fmaxf(sum, 0).  Edge range is
 * [row_ptr[row], row_ptr[row] + num_neighbors[row]) clipped by row_ptr[row+1]
 * and total_nnz (.cu:41-42).  float arithmetic like the kernel, sequential order.
 */
void orc_gcn_fused_forward(const int32_t* row_ptr, const int32_t* col_idx, const float* values,
                           const float* X, const float* W, float* H, const int32_t* num_neighbors,
                           int N, int F_padded, int actual_F, int H_dim, int total_nnz) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int row = 0; row < N; ++row) {
        const int start = row_ptr[row];
        const int nnz = num_neighbors[row];
        for (int j = 0; j < H_dim; ++j) {
            float sum = 0.0f;
            for (int idx = start; idx < start + nnz; ++idx) {
                if (idx < row_ptr[row + 1] && idx < total_nnz) {
                    const int col = col_idx[idx];
                    const float a_val = values[idx];
                    float z = 0.0f;
                    if (col < N)
                        for (int f = 0; f < actual_F; ++f) z += X[(int64_t)col * F_padded + f] * W[(int64_t)f * H_dim + j];
                    sum += a_val * z;
                }
            }
            H[(int64_t)row * H_dim + j] = fmaxf(sum, 0.0f);
        }
    }
}

/*
 * Multi-head sparse GAT aggregation on a CSR (row = destination i, col = source j).
 * mode 1 restates sparseGatConv.forward dgll/nn/Convolution/gatconv.py:122-139:
 *   edge_e = exp(-leakyrelu(a . [h_i || h_j]));  e_rowsum = sum_j edge_e;
 *   h'_i = (sum_j edge_e h_j) / e_rowsum         (no max subtraction, as written)
 * mode 0 restates gatConv.forward :30-54 on the same edge set:
 *   attention = softmax_j(leakyrelu(Wh1_i + Wh2_j)) over adj[i,j] > 0; h' = attention . Wh
 * el[i,h] = a[:D].Wh_i, er[j,h] = a[D:].Wh_j are passed in (ld_e stride).
 * epi: ELU when concat (gatconv.py:53-54,143-145).  Rows without edges give 0
 * (the reference would produce 0/0 = NaN and trip its own assert, :141).
 */
void orc_gat_forward(const int64_t* row_ptr, const int32_t* col, const float* Wh, int64_t ldw,
                     const float* el, const float* er, int64_t ld_e, float* out, int64_t ldo,
                     int64_t n_dst, int heads, int D, float slope, int mode, int epi) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n_dst; ++i) {
        const int64_t b = row_ptr[i], e = row_ptr[i + 1];
        for (int h = 0; h < heads; ++h) {
            double mx = -INFINITY;
            if (mode == 0) {
                for (int64_t k = b; k < e; ++k) {
                    double z = (double)el[i * ld_e + h] + (double)er[(int64_t)col[k] * ld_e + h];
                    z = z > 0 ? z : slope * z;
                    if (z > mx) mx = z;
                }
            } else {
                mx = 0.0; /* reference: plain exp(-z) */
            }
            double rowsum = 0.0;
            double* acc = (double*)calloc((size_t)D, sizeof(double));
            for (int64_t k = b; k < e; ++k) {
                const int64_t j = col[k];
                double z = (double)el[i * ld_e + h] + (double)er[j * ld_e + h];
                z = z > 0 ? z : slope * z;
                const double w = mode == 0 ? exp(z - mx) : exp(-z);
                rowsum += w;
                for (int d = 0; d < D; ++d) acc[d] += w * (double)Wh[j * ldw + (int64_t)h * D + d];
            }
            for (int d = 0; d < D; ++d) {
                const double v = (e > b) ? acc[d] / rowsum : 0.0;
                out[i * ldo + (int64_t)h * D + d] = orc_epi(v, epi);
            }
            free(acc);
        }
    }
}

/*
 * SDDMM: out_e[k] = <A[row(k)], B[col[k]]> — what SpecialSpmmFunction.backward computes
 * as grad_a_dense.view(-1)[edge_idx] after a dense N x N matmul (gatconv.py:76-78).
 */
void orc_sddmm_csr(const int64_t* row_ptr, const int32_t* col, const float* A, int64_t lda,
                   const float* B, int64_t ldb, float* out_e, int64_t n_rows, int F) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n_rows; ++i)
        for (int64_t k = row_ptr[i]; k < row_ptr[i + 1]; ++k) {
            double s = 0.0;
            for (int f = 0; f < F; ++f) s += (double)A[i * lda + f] * (double)B[(int64_t)col[k] * ldb + f];
            out_e[k] = (float)s;
        }
}

/*
 * Row gather: out[i] = table[ids[i]] (features[nodes], dgll/data/dgraph.py:105) and the
 * GraphCacheServer split (dgll/FeatureCache/storage.py:173-194):
 *   gpu_flag[id] ? cache[localid2cacheid[id]] : host[nid_map ? nid_map[id] : id].
 * Byte-exact.  Returns the number of misses (storage.py:213-215).
 */
int64_t orc_gather_rows(const char* table, int64_t stride, const char* host_table, int64_t host_stride,
                        const int64_t* ids, const uint8_t* gpu_flag, const int64_t* local2cache,
                        const int64_t* nid_map, char* out, int64_t out_stride, int64_t n_rows,
                        int64_t row_bytes) {
    int64_t miss = 0;
    for (int64_t i = 0; i < n_rows; ++i) {
        const int64_t id = ids[i];
        const char* src;
        if (gpu_flag) {
            if (gpu_flag[id]) src = table + local2cache[id] * stride;
            else { src = host_table + (nid_map ? nid_map[id] : id) * host_stride; ++miss; }
        } else {
            src = table + id * stride;
        }
        memcpy(out + i * out_stride, src, (size_t)row_bytes);
    }
    return miss;
}

/*
 * Binarized aggregation (no reference code; semantics SURVEY.md §8 a18):
 * packed bit f of row r = (X[r,f] >= 0); cnt[i,f] = sum_{j in N(i)} bit.
 */
void orc_binarize_pack(const float* X, int64_t ldx, uint32_t* packed, int64_t wpr, int64_t n_rows, int F) {
    for (int64_t r = 0; r < n_rows; ++r) {
        for (int64_t w = 0; w < wpr; ++w) packed[r * wpr + w] = 0u;
        for (int f = 0; f < F; ++f)
            if (X[r * ldx + f] >= 0.0f) packed[r * wpr + f / 32] |= 1u << (f % 32);
    }
}

void orc_bin_spmm_counts(const int64_t* row_ptr, const int32_t* col, const uint32_t* packed, int64_t wpr,
                         int32_t* cnt, int64_t n_dst, int F) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n_dst; ++i) {
        for (int f = 0; f < F; ++f) cnt[i * (int64_t)F + f] = 0;
        for (int64_t k = row_ptr[i]; k < row_ptr[i + 1]; ++k) {
            const uint32_t* row = packed + (int64_t)col[k] * wpr;
            for (int f = 0; f < F; ++f) cnt[i * (int64_t)F + f] += (row[f / 32] >> (f % 32)) & 1u;
        }
    }
}

/* C = A[M,K] . B[K,N] (+bias), row-major, double accumulation: torch.mm(x, W)
 * gcnconv.py:30, gcn_model.py:70, gatconv.py:117. */
void orc_gemm(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int64_t M,
              int64_t N, int64_t K, const float* bias, int epi) {
#pragma omp parallel for schedule(static)
    for (int64_t m = 0; m < M; ++m) {
        double* acc = (double*)calloc((size_t)N, sizeof(double));
        for (int64_t k = 0; k < K; ++k) {
            const double a = (double)A[m * lda + k];
            for (int64_t n = 0; n < N; ++n) acc[n] += a * (double)B[k * ldb + n];
        }
        for (int64_t n = 0; n < N; ++n) C[m * ldc + n] = orc_epi(acc[n] + (bias ? (double)bias[n] : 0.0), epi);
        free(acc);
    }
}
