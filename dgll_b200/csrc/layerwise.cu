// layerwise.cu — layer-wise importance sampling (FastGCN / LADIES, "flat" and "WRS" switches) on the device.
//
// Counterpart of the host scipy pipeline in the reference's GPU-Accelerator scripts:
//     Q      = lap_matrix[prev_nodes, :]                          MQLadies.py:78      -> dgllb_csr_slice_rows_*
//     prob_i = sum(Q.multiply(Q), axis=0) (sqrt if flat)          MQLadies.py:79-81   -> dgllb_col_sqsum
//     picks  = np.random.choice(n, s_num, p=prob, replace=False)  utils.py:201        -> dgllb_weighted_choice
//     w      = estWRS_weights / 1/prob/s_num                      utils.py:199-213, MQFastGCN.py:82 -> dgllb_importance_scale
//     adj    = Q[:, picks].multiply(w).tocsr()                    MQLadies.py:84      -> dgllb_csr_select_cols
// All of it is integer / fp64 bookkeeping bound by HBM latency and by the radix sorts, not by arithmetic; the point
// of running it here is that the block lands in HBM next to the feature table, stream-ordered with the layer kernels,
// with two small read-backs per layer (the sliced nnz, then {s_num, block nnz}) instead of a scipy round trip.
//
// Exactness (checked against oracle/layerwise.py): slicing, column selection, local relabelling and the per-row order
// of the emitted CSR are bit-exact; the column sums are accumulated in the same (row, position) order scipy's CSR
// mat-vec uses and each product / sum is rounded separately (no FMA contraction), so prob_i is bit-identical too; the
// WRS recurrence runs the same operations in the same order.  Only the normaliser sum(prob_i) is a tree reduction
// (numpy's is pairwise): probabilities and weights agree to ~1e-15 relative.
// The draw itself uses exponential keys (-log(u)/p, smallest first — the order statistics of successive weighted draws
// without replacement, which is what choice(replace=False, p=) realises) with a counter-based generator keyed by
// (seed, node id): reproducible, independent of the candidate order, but not numpy's MT19937 stream.
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_segmented_sort.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

namespace dgllb {

namespace {

constexpr int kTB = 256;

inline unsigned grid_for(long long n, const DevInfo& di) {
    long long b = (n + kTB - 1) / kTB;
    const long long cap = static_cast<long long>(di.sm_count) * 16;
    if (b > cap) b = cap;
    return static_cast<unsigned>(b < 1 ? 1 : b);
}

__device__ __forceinline__ long long lw_rp(const void* rp, int is64, long long i) {
    return is64 ? static_cast<const long long*>(rp)[i] : static_cast<long long>(static_cast<const int*>(rp)[i]);
}

__device__ __forceinline__ uint64_t lw_mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// ---------------------------------------------------------------- row slice --
__global__ void slice_deg_kernel(const void* __restrict__ rp, int is64, const long long* __restrict__ rows,
                                 long long n_rows, long long* __restrict__ deg) {
    for (long long k = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; k < n_rows;
         k += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = rows[k];
        deg[k] = lw_rp(rp, is64, r + 1) - lw_rp(rp, is64, r);
    }
}

// one warp per sliced row: contiguous copy of its (col, val) run
__global__ void slice_fill_kernel(const void* __restrict__ rp, int is64, const int* __restrict__ col,
                                  const double* __restrict__ vals, const long long* __restrict__ rows,
                                  long long n_rows, const long long* __restrict__ out_rp, int* __restrict__ out_col,
                                  double* __restrict__ out_vals) {
    const int lane = threadIdx.x & 31;
    const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    for (long long k = warp; k < n_rows; k += n_warps) {
        const long long r = rows[k];
        const long long b = lw_rp(rp, is64, r), e = lw_rp(rp, is64, r + 1), o = out_rp[k];
        for (long long j = b + lane; j < e; j += 32) {
            out_col[o + (j - b)] = col[j];
            if (vals) out_vals[o + (j - b)] = vals[j];
        }
    }
}

// ------------------------------------------------------------ column sq-sum --
__global__ void iota_kernel(int* __restrict__ a, long long n) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        a[i] = static_cast<int>(i);
}

__global__ void head_flag_kernel(const int* __restrict__ sorted_col, long long n, int* __restrict__ flag) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        flag[i] = (i == 0 || sorted_col[i] != sorted_col[i - 1]) ? 1 : 0;
}

// thread at a run head sums its column's squares in stored (row-major) order — the order scipy's CSR mat-vec uses;
// product and sum rounded separately, as numpy does
__global__ void col_sum_kernel(const int* __restrict__ sorted_col, const int* __restrict__ sorted_idx,
                               const int* __restrict__ flag, const int* __restrict__ rank,
                               const double* __restrict__ vals, long long n, int flat, int* __restrict__ cand_cols,
                               double* __restrict__ cand_prob) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        if (!flag[i]) continue;
        const int c = sorted_col[i];
        double s = 0.0;
        for (long long j = i; j < n && sorted_col[j] == c; ++j) {
            const double v = vals[sorted_idx[j]];
            s = __dadd_rn(s, __dmul_rn(v, v));
        }
        cand_cols[rank[i]] = c;
        cand_prob[rank[i]] = flat ? sqrt(s) : s;
    }
}

struct IsPositive {
    __device__ __forceinline__ double operator()(const double& x) const { return x > 0.0 ? 1.0 : 0.0; }
};

// stats = {n_cand, total, n_pos}; then prob /= total in place
__global__ void sqsum_stats_kernel(const int* __restrict__ flag, const int* __restrict__ rank, long long n,
                                   const double* __restrict__ total, const double* __restrict__ n_pos,
                                   double* __restrict__ stats) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        stats[0] = n > 0 ? static_cast<double>(rank[n - 1] + flag[n - 1]) : 0.0;
        stats[1] = *total;
        stats[2] = *n_pos;
    }
}

__global__ void normalize_kernel(double* __restrict__ p, long long n, const double* __restrict__ total) {
    const double t = *total;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x)
        p[i] = p[i] / t;
}

// ------------------------------------------------------------ weighted draw --
__global__ void choice_key_kernel(const int* __restrict__ cand_cols, const double* __restrict__ prob, long long n,
                                  unsigned long long seed, unsigned long long* __restrict__ keys,
                                  int* __restrict__ idx) {
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const double p = prob[i];
        unsigned long long kb = 0x7FF0000000000000ull;  // +inf: never drawn
        if (p > 0.0) {
            const long long node = cand_cols ? cand_cols[i] : i;
            const uint64_t r = lw_mix64(lw_mix64(seed) ^ (static_cast<uint64_t>(node) * 0xD1B54A32D192ED03ull));
            const double u = (static_cast<double>(r >> 11) + 1.0) * 0x1.0p-53;  // (0, 1]
            double key = -log(u) / p;
            if (!(key < 1.0e300)) key = 1.0e300;
            kb = static_cast<unsigned long long>(__double_as_longlong(key));
        }
        keys[i] = kb;
        idx[i] = static_cast<int>(i);
    }
}

// the first `fanout` sorted keys that are finite are the draw, in drawing order
__global__ void choice_emit_kernel(const unsigned long long* __restrict__ sorted_keys,
                                   const int* __restrict__ sorted_idx, const int* __restrict__ cand_cols, long long n,
                                   int fanout, int* __restrict__ sel, long long* __restrict__ picks,
                                   long long* __restrict__ count) {
    __shared__ int s_count;
    if (threadIdx.x == 0) s_count = 0;
    __syncthreads();
    int mine = 0;
    for (int k = threadIdx.x; k < fanout; k += blockDim.x) {
        const bool ok = k < n && sorted_keys[k] < 0x7FF0000000000000ull;
        const int i = ok ? sorted_idx[k] : -1;
        sel[k] = i;
        picks[k] = ok ? (cand_cols ? static_cast<long long>(cand_cols[i]) : static_cast<long long>(i)) : -1ll;
        mine += ok;
    }
    atomicAdd(&s_count, mine);
    __syncthreads();
    if (threadIdx.x == 0) *count = s_count;
}

// ------------------------------------------------------- importance weights --
// mode 0: 1/p/m (MQFastGCN.py:82).  mode 1: the WRS estimator utils.py:199-213 —
//   for i in 0..m-1: alpha = n/(i+1)/(n-i); w[i] = (1-p_sum)/p_i*alpha; w[:i] = w[:i]*(1-alpha)+alpha; p_sum += p_i
// restated per element: w_i starts at step i and is then updated by every later alpha, same operations, same order.
__global__ void importance_scale_kernel(const double* __restrict__ prob, const int* __restrict__ sel,
                                        const long long* __restrict__ count, int cap, long long n_total, int mode,
                                        double* __restrict__ alpha, double* __restrict__ psum,
                                        double* __restrict__ scale) {
    const int m = static_cast<int>(min(static_cast<long long>(cap), *count));
    if (mode == 0) {
        for (int i = threadIdx.x; i < cap; i += blockDim.x)
            scale[i] = i < m ? (1.0 / prob[sel[i]]) / static_cast<double>(m) : 0.0;
        return;
    }
    for (int i = threadIdx.x; i < m; i += blockDim.x)
        alpha[i] = (static_cast<double>(n_total) / static_cast<double>(i + 1)) / static_cast<double>(n_total - i);
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < m; ++i) {
            psum[i] = s;
            s = __dadd_rn(s, prob[sel[i]]);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cap; i += blockDim.x) {
        if (i >= m) {
            scale[i] = 0.0;
            continue;
        }
        double w = __dmul_rn(__ddiv_rn(__dsub_rn(1.0, psum[i]), prob[sel[i]]), alpha[i]);
        for (int t = i + 1; t < m; ++t) {
            const double a = alpha[t];
            w = __dadd_rn(__dmul_rn(w, __dsub_rn(1.0, a)), a);
        }
        scale[i] = w;
    }
}

// ---------------------------------------------------------- column selection --
__global__ void scatter_pos_kernel(int* __restrict__ pos, const long long* __restrict__ picks,
                                   const long long* __restrict__ count, long long cap, int reset) {
    const long long m = count ? min(cap, *count) : cap;
    for (long long k = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; k < m;
         k += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long c = picks[k];
        if (c >= 0) pos[c] = reset ? -1 : static_cast<int>(k);
    }
}

__global__ void select_flag_kernel(const int* __restrict__ q_col, const long long* __restrict__ q_rp,
                                   long long n_rows, long long cap, const int* __restrict__ pos,
                                   int* __restrict__ flag) {
    const long long nnz = q_rp[n_rows];
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < cap;
         e += static_cast<long long>(gridDim.x) * blockDim.x)
        flag[e] = (e < nnz && pos[q_col[e]] >= 0) ? 1 : 0;
}

__global__ void select_ptr_kernel(const long long* __restrict__ q_rp, long long n_rows, long long cap,
                                  const int* __restrict__ flag, const int* __restrict__ rank,
                                  long long* __restrict__ out_rp) {
    for (long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; r <= n_rows;
         r += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long e = q_rp[r];
        out_rp[r] = e < cap ? rank[e] : (cap > 0 ? rank[cap - 1] + flag[cap - 1] : 0);
    }
}

__global__ void select_fill_kernel(const int* __restrict__ q_col, const double* __restrict__ q_vals,
                                   const long long* __restrict__ q_rp, long long n_rows, const int* __restrict__ pos,
                                   const double* __restrict__ scale, const int* __restrict__ flag,
                                   const int* __restrict__ rank, int* __restrict__ out_col,
                                   double* __restrict__ out_vals) {
    const long long nnz = q_rp[n_rows];
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < nnz;
         e += static_cast<long long>(gridDim.x) * blockDim.x) {
        if (!flag[e]) continue;
        const int k = pos[q_col[e]];
        out_col[rank[e]] = k;
        if (out_vals) {
            const double v = q_vals ? q_vals[e] : 1.0;
            out_vals[rank[e]] = scale ? __dmul_rn(v, scale[k]) : v;
        }
    }
}

struct Workspace {
    char* base = nullptr;
    size_t used = 0, size = 0;
    cudaStream_t st;
    explicit Workspace(cudaStream_t s) : st(s) {}
    static size_t al(size_t b) { return (b + 255) & ~static_cast<size_t>(255); }
    size_t reserve(size_t b) {
        const size_t o = size;
        size += al(b);
        return o;
    }
    cudaError_t commit() { return size ? cudaMallocAsync(&base, size, st) : cudaSuccess; }
    template <typename T>
    T* at(size_t off) const { return reinterpret_cast<T*>(base + off); }
    ~Workspace() {
        if (base) cudaFreeAsync(base, st);
    }
};

int bits_for(long long n) {
    int b = 1;
    while (b < 32 && (1ll << b) < n) ++b;
    return b;
}

}  // namespace
}  // namespace dgllb

using namespace dgllb;

#define LW_CUB(expr, what)                                                     \
    do {                                                                       \
        cudaError_t e__ = (expr);                                              \
        if (e__ != cudaSuccess) {                                              \
            set_error("%s: %s", what, cudaGetErrorString(e__));                \
            return DGLLB_ERR_CUDA;                                             \
        }                                                                      \
    } while (0)

extern "C" int dgllb_csr_slice_rows_ptr(const void* row_ptr, int row_ptr_is64, const int64_t* rows, int64_t n_rows,
                                        int64_t* out_row_ptr, void* stream) {
    DGLLB_REQUIRE(n_rows >= 0, "csr_slice_rows: negative size");
    DGLLB_REQUIRE(out_row_ptr && (n_rows == 0 || (row_ptr && rows)), "csr_slice_rows: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DevInfo di;
    int rc = get_devinfo(&di);
    if (rc != DGLLB_OK) return rc;
    if (n_rows == 0) {
        DGLLB_CUDA_TRY(cudaMemsetAsync(out_row_ptr, 0, sizeof(int64_t), st));
        return DGLLB_OK;
    }
    Workspace ws(st);
    size_t cub_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, static_cast<long long*>(nullptr), static_cast<long long*>(nullptr),
                                  static_cast<int>(n_rows + 1), st);
    const size_t o_deg = ws.reserve(sizeof(long long) * (n_rows + 1)), o_cub = ws.reserve(cub_bytes);
    LW_CUB(ws.commit(), "csr_slice_rows");
    long long* deg = ws.at<long long>(o_deg);
    DGLLB_CUDA_TRY(cudaMemsetAsync(deg + n_rows, 0, sizeof(long long), st));
    slice_deg_kernel<<<grid_for(n_rows, di), kTB, 0, st>>>(row_ptr, row_ptr_is64,
                                                           reinterpret_cast<const long long*>(rows), n_rows, deg);
    DGLLB_LAUNCH_CHECK();
    LW_CUB(cub::DeviceScan::ExclusiveSum(ws.at<void>(o_cub), cub_bytes, deg, reinterpret_cast<long long*>(out_row_ptr),
                                         static_cast<int>(n_rows + 1), st),
           "csr_slice_rows: scan");
    g_launch_count.fetch_add(2);
    return DGLLB_OK;
}

extern "C" int dgllb_csr_slice_rows_fill(const void* row_ptr, int row_ptr_is64, const int32_t* col_idx,
                                         const double* values, const int64_t* rows, int64_t n_rows,
                                         const int64_t* out_row_ptr, int32_t* out_col, double* out_values,
                                         void* stream) {
    DGLLB_REQUIRE(n_rows >= 0, "csr_slice_rows: negative size");
    if (n_rows == 0) return DGLLB_OK;
    DGLLB_REQUIRE(row_ptr && col_idx && rows && out_row_ptr && out_col && (!values || out_values),
                  "csr_slice_rows: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DevInfo di;
    int rc = get_devinfo(&di);
    if (rc != DGLLB_OK) return rc;
    slice_fill_kernel<<<grid_for(n_rows * 32, di), kTB, 0, st>>>(
        row_ptr, row_ptr_is64, col_idx, values, reinterpret_cast<const long long*>(rows), n_rows,
        reinterpret_cast<const long long*>(out_row_ptr), out_col, out_values);
    DGLLB_LAUNCH_CHECK();
    g_launch_count.fetch_add(1);
    return DGLLB_OK;
}

extern "C" int dgllb_col_sqsum(const int32_t* col_idx, const double* values, int64_t nnz, int64_t n_cols, int flat,
                               int32_t* cand_cols, double* cand_prob, double* stats, void* stream) {
    DGLLB_REQUIRE(nnz >= 0 && n_cols >= 0 && nnz < (1ll << 31), "col_sqsum: bad size");
    DGLLB_REQUIRE(stats && (nnz == 0 || (col_idx && values && cand_cols && cand_prob)), "col_sqsum: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DevInfo di;
    int rc = get_devinfo(&di);
    if (rc != DGLLB_OK) return rc;
    if (nnz == 0) {
        DGLLB_CUDA_TRY(cudaMemsetAsync(stats, 0, 3 * sizeof(double), st));
        return DGLLB_OK;
    }
    const int n = static_cast<int>(nnz);
    const int end_bit = bits_for(n_cols > 1 ? n_cols : 2);
    Workspace ws(st);
    size_t b_sort = 0, b_scan = 0, b_red = 0, b_red2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, b_sort, static_cast<const int*>(nullptr), static_cast<int*>(nullptr),
                                    static_cast<const int*>(nullptr), static_cast<int*>(nullptr), n, 0, end_bit, st);
    cub::DeviceScan::ExclusiveSum(nullptr, b_scan, static_cast<int*>(nullptr), static_cast<int*>(nullptr), n, st);
    cub::DeviceReduce::Sum(nullptr, b_red, static_cast<double*>(nullptr), static_cast<double*>(nullptr), n, st);
    cub::TransformInputIterator<double, IsPositive, const double*> pos_it(cand_prob, IsPositive());
    cub::DeviceReduce::Sum(nullptr, b_red2, pos_it, static_cast<double*>(nullptr), n, st);
    size_t b_cub = b_sort;
    if (b_scan > b_cub) b_cub = b_scan;
    if (b_red > b_cub) b_cub = b_red;
    if (b_red2 > b_cub) b_cub = b_red2;
    const size_t o_iota = ws.reserve(sizeof(int) * nnz), o_scol = ws.reserve(sizeof(int) * nnz),
                 o_sidx = ws.reserve(sizeof(int) * nnz), o_flag = ws.reserve(sizeof(int) * nnz),
                 o_rank = ws.reserve(sizeof(int) * nnz), o_tot = ws.reserve(2 * sizeof(double)),
                 o_cub = ws.reserve(b_cub);
    LW_CUB(ws.commit(), "col_sqsum");
    int *iota = ws.at<int>(o_iota), *scol = ws.at<int>(o_scol), *sidx = ws.at<int>(o_sidx), *flag = ws.at<int>(o_flag),
        *rank = ws.at<int>(o_rank);
    double* tot = ws.at<double>(o_tot);
    const unsigned g = grid_for(nnz, di);
    iota_kernel<<<g, kTB, 0, st>>>(iota, nnz);
    LW_CUB(cub::DeviceRadixSort::SortPairs(ws.at<void>(o_cub), b_sort, col_idx, scol, iota, sidx, n, 0, end_bit, st),
           "col_sqsum: sort");
    head_flag_kernel<<<g, kTB, 0, st>>>(scol, nnz, flag);
    LW_CUB(cub::DeviceScan::ExclusiveSum(ws.at<void>(o_cub), b_scan, flag, rank, n, st), "col_sqsum: scan");
    DGLLB_CUDA_TRY(cudaMemsetAsync(cand_prob, 0, sizeof(double) * nnz, st));
    DGLLB_CUDA_TRY(cudaMemsetAsync(cand_cols, 0xFF, sizeof(int) * nnz, st));
    col_sum_kernel<<<g, kTB, 0, st>>>(scol, sidx, flag, rank, values, nnz, flat, cand_cols, cand_prob);
    LW_CUB(cub::DeviceReduce::Sum(ws.at<void>(o_cub), b_red, cand_prob, tot, n, st), "col_sqsum: reduce");
    LW_CUB(cub::DeviceReduce::Sum(ws.at<void>(o_cub), b_red2, pos_it, tot + 1, n, st), "col_sqsum: count");
    sqsum_stats_kernel<<<1, 32, 0, st>>>(flag, rank, nnz, tot, tot + 1, stats);
    normalize_kernel<<<g, kTB, 0, st>>>(cand_prob, nnz, tot);
    DGLLB_LAUNCH_CHECK();
    g_launch_count.fetch_add(9);
    return DGLLB_OK;
}

extern "C" int dgllb_weighted_choice(const int32_t* cand_cols, const double* cand_prob, int64_t n_cand, int fanout,
                                     uint64_t seed, int32_t* sel, int64_t* picks, int64_t* count, void* stream) {
    DGLLB_REQUIRE(n_cand >= 0 && n_cand < (1ll << 31) && fanout >= 0, "weighted_choice: bad size");
    DGLLB_REQUIRE(count && (fanout == 0 || (sel && picks)) && (n_cand == 0 || cand_prob), "weighted_choice: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DevInfo di;
    int rc = get_devinfo(&di);
    if (rc != DGLLB_OK) return rc;
    const int n = static_cast<int>(n_cand);
    Workspace ws(st);
    size_t b_sort = 0;
    if (n > 0)
        cub::DeviceRadixSort::SortPairs(nullptr, b_sort, static_cast<const unsigned long long*>(nullptr),
                                        static_cast<unsigned long long*>(nullptr), static_cast<const int*>(nullptr),
                                        static_cast<int*>(nullptr), n, 0, 64, st);
    const size_t nn = n > 0 ? n : 1;
    const size_t o_k = ws.reserve(8 * nn), o_sk = ws.reserve(8 * nn), o_i = ws.reserve(4 * nn), o_si = ws.reserve(4 * nn),
                 o_cub = ws.reserve(b_sort);
    LW_CUB(ws.commit(), "weighted_choice");
    unsigned long long *keys = ws.at<unsigned long long>(o_k), *skeys = ws.at<unsigned long long>(o_sk);
    int *idx = ws.at<int>(o_i), *sidx = ws.at<int>(o_si);
    if (n > 0) {
        choice_key_kernel<<<grid_for(n, di), kTB, 0, st>>>(cand_cols, cand_prob, n, seed, keys, idx);
        LW_CUB(cub::DeviceRadixSort::SortPairs(ws.at<void>(o_cub), b_sort, keys, skeys, idx, sidx, n, 0, 64, st),
               "weighted_choice: sort");
    }
    choice_emit_kernel<<<1, kTB, 0, st>>>(skeys, sidx, cand_cols, n, fanout, sel, reinterpret_cast<long long*>(picks),
                                          reinterpret_cast<long long*>(count));
    DGLLB_LAUNCH_CHECK();
    g_launch_count.fetch_add(3);
    return DGLLB_OK;
}

extern "C" int dgllb_importance_scale(const double* cand_prob, const int32_t* sel, const int64_t* count, int cap,
                                      int64_t n_total, int mode, double* scale, void* stream) {
    DGLLB_REQUIRE(cap >= 0 && n_total >= 0 && (mode == 0 || mode == 1), "importance_scale: bad argument");
    if (cap == 0) return DGLLB_OK;
    DGLLB_REQUIRE(cand_prob && sel && count && scale, "importance_scale: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Workspace ws(st);
    const size_t o_a = ws.reserve(sizeof(double) * cap), o_p = ws.reserve(sizeof(double) * cap);
    LW_CUB(ws.commit(), "importance_scale");
    importance_scale_kernel<<<1, 1024, 0, st>>>(cand_prob, sel, reinterpret_cast<const long long*>(count), cap, n_total,
                                                mode, ws.at<double>(o_a), ws.at<double>(o_p), scale);
    DGLLB_LAUNCH_CHECK();
    g_launch_count.fetch_add(1);
    return DGLLB_OK;
}

extern "C" int dgllb_scatter_pos(int32_t* pos, const int64_t* picks, const int64_t* count, int64_t cap, int reset,
                                 void* stream) {
    DGLLB_REQUIRE(cap >= 0, "scatter_pos: negative size");
    if (cap == 0) return DGLLB_OK;
    DGLLB_REQUIRE(pos && picks, "scatter_pos: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DevInfo di;
    int rc = get_devinfo(&di);
    if (rc != DGLLB_OK) return rc;
    scatter_pos_kernel<<<grid_for(cap, di), kTB, 0, st>>>(pos, reinterpret_cast<const long long*>(picks),
                                                          reinterpret_cast<const long long*>(count), cap, reset);
    DGLLB_LAUNCH_CHECK();
    g_launch_count.fetch_add(1);
    return DGLLB_OK;
}

extern "C" int dgllb_csr_select_cols(const int64_t* q_row_ptr, const int32_t* q_col, const double* q_values,
                                     int64_t n_rows, int64_t nnz_cap, const int32_t* pos, const double* scale,
                                     int64_t* out_row_ptr, int32_t* out_col, double* out_values, void* stream) {
    DGLLB_REQUIRE(n_rows >= 0 && nnz_cap >= 0 && nnz_cap < (1ll << 31), "csr_select_cols: bad size");
    DGLLB_REQUIRE(q_row_ptr && out_row_ptr && (nnz_cap == 0 || (q_col && pos && out_col)), "csr_select_cols: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DevInfo di;
    int rc = get_devinfo(&di);
    if (rc != DGLLB_OK) return rc;
    if (nnz_cap == 0) {
        DGLLB_CUDA_TRY(cudaMemsetAsync(out_row_ptr, 0, sizeof(int64_t) * (n_rows + 1), st));
        return DGLLB_OK;
    }
    const int n = static_cast<int>(nnz_cap);
    const long long* q_rp = reinterpret_cast<const long long*>(q_row_ptr);
    long long* o_rp = reinterpret_cast<long long*>(out_row_ptr);
    Workspace ws(st);
    size_t b_scan = 0, b_seg = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b_scan, static_cast<int*>(nullptr), static_cast<int*>(nullptr), n, st);
    if (out_values)
        cub::DeviceSegmentedSort::SortPairs(nullptr, b_seg, static_cast<const int*>(nullptr), static_cast<int*>(nullptr),
                                            static_cast<const double*>(nullptr), static_cast<double*>(nullptr), n,
                                            static_cast<int>(n_rows), o_rp, o_rp + 1, st);
    else
        cub::DeviceSegmentedSort::SortKeys(nullptr, b_seg, static_cast<const int*>(nullptr), static_cast<int*>(nullptr),
                                           n, static_cast<int>(n_rows), o_rp, o_rp + 1, st);
    const size_t o_flag = ws.reserve(sizeof(int) * nnz_cap), o_rank = ws.reserve(sizeof(int) * nnz_cap),
                 o_tc = ws.reserve(sizeof(int) * nnz_cap), o_tv = ws.reserve(sizeof(double) * nnz_cap),
                 o_cub = ws.reserve(b_scan > b_seg ? b_scan : b_seg);
    LW_CUB(ws.commit(), "csr_select_cols");
    int *flag = ws.at<int>(o_flag), *rank = ws.at<int>(o_rank), *tc = ws.at<int>(o_tc);
    double* tv = out_values ? ws.at<double>(o_tv) : nullptr;
    const unsigned g = grid_for(nnz_cap, di);
    select_flag_kernel<<<g, kTB, 0, st>>>(q_col, q_rp, n_rows, nnz_cap, pos, flag);
    LW_CUB(cub::DeviceScan::ExclusiveSum(ws.at<void>(o_cub), b_scan, flag, rank, n, st), "csr_select_cols: scan");
    select_ptr_kernel<<<grid_for(n_rows + 1, di), kTB, 0, st>>>(q_rp, n_rows, nnz_cap, flag, rank, o_rp);
    select_fill_kernel<<<g, kTB, 0, st>>>(q_col, q_values, q_rp, n_rows, pos, scale, flag, rank, tc, tv);
    DGLLB_LAUNCH_CHECK();
    if (n_rows > 0) {
        if (out_values)
            LW_CUB(cub::DeviceSegmentedSort::SortPairs(ws.at<void>(o_cub), b_seg, tc, out_col, tv, out_values, n,
                                                       static_cast<int>(n_rows), o_rp, o_rp + 1, st),
                   "csr_select_cols: segmented sort");
        else
            LW_CUB(cub::DeviceSegmentedSort::SortKeys(ws.at<void>(o_cub), b_seg, tc, out_col, n,
                                                      static_cast<int>(n_rows), o_rp, o_rp + 1, st),
                   "csr_select_cols: segmented sort");
    }
    g_launch_count.fetch_add(5);
    return DGLLB_OK;
}
