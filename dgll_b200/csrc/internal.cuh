// internal.cuh — declarations shared between the translation units of the library
// (not part of the C ABI).
#pragma once
#include "common.cuh"

// nnz-split schedule of a static CSR (opaque to C callers, include/dgll_b200.h)
struct dgllb_csr_plan {
    int chunk_edges;
    long long n_rows;
    long long n_heavy_rows;
    long long n_items;
    int* heavy_rows;  // device: rows longer than chunk_edges
    int2* items;      // device: (row, chunk index) for every chunk of every heavy row
};

namespace dgllb {

struct SpmmParams {
    const void* row_ptr;
    int rp64;
    const int* col;
    const float* vals;
    const void* X;
    long long ldx;
    float* out;
    long long ldo;
    long long n_dst;
    int F;
    int n_slabs;
    int mean;
    const float* row_scale;
    const float* addend;
    long long ld_add;
    const float* bias;
    int epi;
    int* argmax;
    // nnz-split (heavy rows)
    const int2* heavy_items;  // (row, chunk)
    long long n_heavy_items;
    int chunk_edges;          // 0 = no plan
    float* heavy_ws;          // [n_heavy_items, ld_hws] partial rows of the split rows (merged in chunk order)
    long long ld_hws;
    const int* row_cnt;       // optional per-row edge-count clamp (legacy num_neighbors)
    long long nnz_hint;       // host-known upper bound on row_ptr[n_dst]; < 0 = unknown
    int out_vec;              // out rows are 16-byte aligned: 128-bit stores / RED.ADD.128 allowed
};

// run the aggregation kernel(s) for a filled parameter block (spmm.cu)
int spmm_run(SpmmParams& p, int x_dtype, bool is_max, const dgllb_csr_plan* plan, cudaStream_t st);

// whole-row rolling-window aggregation for short rows (spmm_rows.cu); DGLLB_ERR_UNSUPPORTED = shape not taken
int spmm_rows_try(const SpmmParams& p, int x_dtype, cudaStream_t st);

// exact-fp32 SIMT GEMM (gemm_simt.cu)
int gemm_simt(const float* A, long long lda, int transA, const float* B, long long ldb, int transB, float* C,
              long long ldc, long long M, long long N, long long K, const float* bias, int epi,
              int accumulate, cudaStream_t st);

// tcgen05 bf16 GEMM (gemm_tcgen05.cu); returns DGLLB_ERR_UNSUPPORTED for shapes it does not take
int gemm_tcgen05(const float* A, long long lda, int transA, const float* B, long long ldb, int transB,
                 float* C, long long ldc, long long M, long long N, long long K, const float* bias, int epi,
                 int accumulate, cudaStream_t st);

// tcgen05 TF32 GEMM straight from the fp32 tensors (gemm_tf32.cu): TMA loads, K-major or MN-major operands, no packing
int gemm_tf32(const float* A, long long lda, int transA, const float* B, long long ldb, int transB, float* C,
              long long ldc, long long M, long long N, long long K, const float* bias, int epi, int accumulate,
              cudaStream_t st);

// 3xTF32: every fp32 operand word split into tf32 hi + lo in shared memory, three MMAs per k-step: fp32-grade results
// (~1e-6 of max|ref|) on the tensor cores (gemm_tf32.cu)
int gemm_tf32x3(const float* A, long long lda, int transA, const float* B, long long ldb, int transB, float* C,
                long long ldc, long long M, long long N, long long K, const float* bias, int epi, int accumulate,
                cudaStream_t st);

}  // namespace dgllb
