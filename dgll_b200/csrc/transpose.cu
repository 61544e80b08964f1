// transpose.cu — CSR transpose for the backward of the aggregation
// (grad_X = A^T grad_out; the reference writes `a.t().matmul(grad_output)`,
// dgll/nn/Convolution/gatconv.py:80, and relies on ATen's COO transpose).
//
// Stable counting sort by column without atomics: expand row ids per edge,
// CUB radix-sort (column, edge) pairs (stable), read the transposed row
// pointers off the sorted keys with a lower-bound search, then permute.
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>

namespace dgllb {

__device__ __forceinline__ long long ld_rp(const void* p, int is64, long long i) {
    return is64 ? reinterpret_cast<const long long*>(p)[i]
                : static_cast<long long>(reinterpret_cast<const int*>(p)[i]);
}

// rowid[e] = row that owns edge e (binary search over row_ptr), eid[e] = e
__global__ void expand_rows_kernel(const void* row_ptr, int rp64, long long n_rows, long long nnz,
                                   int* __restrict__ rowid, int* __restrict__ eid) {
    const long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    long long lo = 0, hi = n_rows;  // find last row with row_ptr[row] <= e
    while (hi - lo > 1) {
        const long long mid = (lo + hi) >> 1;
        if (ld_rp(row_ptr, rp64, mid) <= e) lo = mid; else hi = mid;
    }
    rowid[e] = static_cast<int>(lo);
    eid[e] = static_cast<int>(e);
}

// t_row_ptr[c] = first position in sorted keys with key >= c
__global__ void lower_bound_kernel(const int* __restrict__ keys, long long nnz, long long n_cols,
                                   void* t_row_ptr, int rp64) {
    const long long c = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (c > n_cols) return;
    long long lo = 0, hi = nnz;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (keys[mid] < c) lo = mid + 1; else hi = mid;
    }
    if (rp64) reinterpret_cast<long long*>(t_row_ptr)[c] = lo;
    else reinterpret_cast<int*>(t_row_ptr)[c] = static_cast<int>(lo);
}

__global__ void permute_kernel(const int* __restrict__ perm, const int* __restrict__ rowid,
                               const float* __restrict__ values, long long nnz, int* __restrict__ t_col,
                               float* __restrict__ t_val, int* __restrict__ perm_out) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const int e = perm[i];
    t_col[i] = rowid[e];
    if (t_val) t_val[i] = values ? values[e] : 1.f;
    if (perm_out) perm_out[i] = e;
}

}  // namespace dgllb

using namespace dgllb;

extern "C" int dgllb_csr_transpose(const void* row_ptr, int row_ptr_is64, const int32_t* col_idx,
                                   const float* values, int64_t n_rows, int64_t n_cols, int64_t nnz,
                                   void* t_row_ptr, int32_t* t_col_idx, float* t_values,
                                   int32_t* perm, void* stream) {
    DGLLB_REQUIRE(row_ptr && t_row_ptr, "csr_transpose: null pointer");
    DGLLB_REQUIRE(n_rows >= 0 && n_cols >= 0 && nnz >= 0, "csr_transpose: negative size");
    DGLLB_REQUIRE(nnz < (1ll << 31) - 1 && n_rows < (1ll << 31) && n_cols < (1ll << 31),
                  "csr_transpose: sizes must fit int32 (nnz=%lld)", (long long)nnz);
    { DevInfo di_; int rc_ = get_devinfo(&di_); if (rc_ != DGLLB_OK) return rc_; }  // also configures the workspace pool
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int tb = 256;
    if (nnz == 0) {
        lower_bound_kernel<<<static_cast<unsigned>((n_cols + 1 + tb - 1) / tb), tb, 0, st>>>(
            nullptr, 0, n_cols, t_row_ptr, row_ptr_is64);
        DGLLB_LAUNCH_CHECK();
        return DGLLB_OK;
    }
    DGLLB_REQUIRE(col_idx && t_col_idx, "csr_transpose: null pointer");
    // workspace: rowid, eid, keys_out, eid_out (4 x nnz ints) + CUB temp
    // keys 0..n_cols: a column id equal to n_cols is PADDING — it sorts behind every real column and lands in no row of
    // the result (t_row_ptr[n_cols] = number of real entries), so fixed-capacity edge arrays can be transposed as they are
    int end_bit = 1;
    while ((1ll << end_bit) <= n_cols && end_bit < 32) ++end_bit;
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, static_cast<const int*>(nullptr),
                                    static_cast<int*>(nullptr), static_cast<const int*>(nullptr),
                                    static_cast<int*>(nullptr), static_cast<int>(nnz), 0, end_bit, st);
    const size_t ints = static_cast<size_t>(nnz) * sizeof(int);
    const size_t ints_al = (ints + 255) & ~static_cast<size_t>(255);
    char* ws = nullptr;
    DGLLB_CUDA_TRY(cudaMallocAsync(&ws, 4 * ints_al + cub_bytes, st));
    int* rowid = reinterpret_cast<int*>(ws);
    int* eid = reinterpret_cast<int*>(ws + ints_al);
    int* keys_out = reinterpret_cast<int*>(ws + 2 * ints_al);
    int* eid_out = reinterpret_cast<int*>(ws + 3 * ints_al);
    void* cub_ws = ws + 4 * ints_al;
    const unsigned eg = static_cast<unsigned>((nnz + tb - 1) / tb);
    int rc = DGLLB_OK;
    do {
        expand_rows_kernel<<<eg, tb, 0, st>>>(row_ptr, row_ptr_is64, n_rows, nnz, rowid, eid);
        g_launch_count.fetch_add(1);
        cudaError_t e = cub::DeviceRadixSort::SortPairs(cub_ws, cub_bytes, col_idx, keys_out, eid, eid_out,
                                                        static_cast<int>(nnz), 0, end_bit, st);
        g_launch_count.fetch_add(4);
        if (e != cudaSuccess) { set_error("csr_transpose: cub sort: %s", cudaGetErrorString(e)); rc = DGLLB_ERR_CUDA; break; }
        lower_bound_kernel<<<static_cast<unsigned>((n_cols + 1 + tb - 1) / tb), tb, 0, st>>>(
            keys_out, nnz, n_cols, t_row_ptr, row_ptr_is64);
        g_launch_count.fetch_add(1);
        permute_kernel<<<eg, tb, 0, st>>>(eid_out, rowid, values, nnz, t_col_idx, t_values, perm);
        g_launch_count.fetch_add(1);
        e = cudaGetLastError();
        if (e != cudaSuccess) { set_error("csr_transpose: %s", cudaGetErrorString(e)); rc = DGLLB_ERR_CUDA; }
    } while (0);
    cudaFreeAsync(ws, st);
    return rc;
}
