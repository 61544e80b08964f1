// sampler.cu — uniform neighbour sampling without replacement on the device.
//
// Device-side counterpart of Base_sampler.sample_neighbours
// (dgll/sampling/base_sampler.py:45-58): min(deg, fanout) distinct in-neighbours
// per seed, all neighbours when deg <= fanout (original order), emitted as a CSR
// block by destination with global source ids (the reference's block is the
// same edge list as (src[], dst[]) tensors, base_sampler.py:30-43).
// Bit-exact parity with Python's Mersenne-Twister random.sample is impossible on
// device: the parity path uses host-sampled lists; this
// kernel is validated structurally (subset, count, no duplicates).
//
// One warp per seed.  deg > fanout: Floyd's algorithm draws `fanout` distinct
// positions (lanes hold the chosen set, membership by ballot), positions are
// ranked with shuffles so the row keeps ascending neighbour order.  fanout <= 32.
#include "common.cuh"
#include <cub/device/device_scan.cuh>

namespace dgllb {

__device__ __forceinline__ long long smp_rp(const void* p, int is64, long long i) {
    return is64 ? reinterpret_cast<const long long*>(p)[i]
                : static_cast<long long>(reinterpret_cast<const int*>(p)[i]);
}
__device__ __forceinline__ long long smp_seed(const void* p, int is64, long long i) {
    return is64 ? reinterpret_cast<const long long*>(p)[i]
                : static_cast<long long>(reinterpret_cast<const int*>(p)[i]);
}

// splitmix64: counter-based, one value per (rng_seed, seed index, draw)
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void sample_count_kernel(const void* row_ptr, int rp64, const void* seeds, int s64,
                                    long long n_seeds, int fanout, int* __restrict__ counts) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i > n_seeds) return;
    if (i == n_seeds) { counts[i] = 0; return; }
    const long long v = smp_seed(seeds, s64, i);
    if (v < 0) { counts[i] = 0; return; }  // padding slot of a fixed-capacity seed array
    const long long deg = smp_rp(row_ptr, rp64, v + 1) - smp_rp(row_ptr, rp64, v);
    counts[i] = static_cast<int>(fanout < 0 ? deg : min(deg, static_cast<long long>(fanout)));
}

__global__ void __launch_bounds__(256)
sample_fill_kernel(const void* row_ptr, int rp64, const int* __restrict__ col, const void* seeds, int s64,
                   long long n_seeds, int fanout, uint64_t rng_seed, const unsigned long long* __restrict__ rng_offset,
                   const int* __restrict__ out_row_ptr, int* __restrict__ out_col) {
    const int lane = threadIdx.x & 31;
    const long long i = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (i >= n_seeds) return;
    const long long v = smp_seed(seeds, s64, i);
    if (v < 0) return;                       // padding slot
    if (rng_offset) rng_seed += *rng_offset; // per-replay offset kept in device memory (CUDA-graph friendly)
    const long long beg = smp_rp(row_ptr, rp64, v), deg = smp_rp(row_ptr, rp64, v + 1) - beg;
    const long long ob = out_row_ptr[i];
    if (fanout < 0 || deg <= fanout) {
        for (long long k = lane; k < deg; k += 32) out_col[ob + k] = __ldg(col + beg + k);
        return;
    }
    // Floyd: for j = deg-fanout .. deg-1: t = U[0, j]; pick t unless already chosen, else pick j.
    long long mine = -1;  // lane s holds the s-th chosen position
    const uint64_t base = mix64(rng_seed ^ (static_cast<uint64_t>(i) * 0xD1B54A32D192ED03ull));
    for (int s = 0; s < fanout; ++s) {
        const long long j = deg - fanout + s;
        const uint64_t r = mix64(base + static_cast<uint64_t>(s));
        // unbiased enough for sampling: 64-bit multiply-high range reduction
        long long t = static_cast<long long>(__umul64hi(r, static_cast<uint64_t>(j + 1)));
        const unsigned hit = __ballot_sync(0xffffffffu, mine == t);
        if (hit) t = j;
        if (lane == s) mine = t;
    }
    // rank the chosen positions (ascending) so the row keeps neighbour order
    int rank = 0;
    for (int s = 0; s < fanout; ++s) {
        const long long other = __shfl_sync(0xffffffffu, mine, s);
        if (lane < fanout && other < mine) ++rank;
    }
    if (lane < fanout) out_col[ob + rank] = __ldg(col + beg + mine);
}

}  // namespace dgllb

using namespace dgllb;

static int sample_neighbors_impl(const void* row_ptr, int row_ptr_is64, const int32_t* col_idx, const void* seeds,
                                 int seeds_is64, int64_t n_seeds, int fanout, uint64_t rng_seed,
                                 const unsigned long long* rng_offset, int32_t* out_row_ptr, int32_t* out_col,
                                 void* stream) {
    DGLLB_REQUIRE(n_seeds >= 0, "sample_neighbors: negative n_seeds");
    DGLLB_REQUIRE(row_ptr && out_row_ptr && (n_seeds == 0 || seeds), "sample_neighbors: null pointer");
    if (fanout > 32) {
        set_error("sample_neighbors: fanout %d > 32 not supported by this build", fanout);
        return DGLLB_ERR_UNSUPPORTED;
    }
    { DevInfo di_; int rc_ = get_devinfo(&di_); if (rc_ != DGLLB_OK) return rc_; }  // also configures the workspace pool
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int tb = 256;
    int* counts = nullptr;
    size_t cub_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, static_cast<int*>(nullptr), static_cast<int*>(nullptr),
                                  static_cast<int>(n_seeds + 1), st);
    const size_t cnt_bytes = ((n_seeds + 1) * sizeof(int) + 255) & ~static_cast<size_t>(255);
    char* ws = nullptr;
    DGLLB_CUDA_TRY(cudaMallocAsync(&ws, cnt_bytes + cub_bytes, st));
    counts = reinterpret_cast<int*>(ws);
    int rc = DGLLB_OK;
    do {
        sample_count_kernel<<<static_cast<unsigned>((n_seeds + 1 + tb - 1) / tb), tb, 0, st>>>(
            row_ptr, row_ptr_is64, seeds, seeds_is64, n_seeds, fanout, counts);
        g_launch_count.fetch_add(1);
        cudaError_t e = cub::DeviceScan::ExclusiveSum(ws + cnt_bytes, cub_bytes, counts, out_row_ptr,
                                                      static_cast<int>(n_seeds + 1), st);
        g_launch_count.fetch_add(1);
        if (e != cudaSuccess) { set_error("sample_neighbors: scan: %s", cudaGetErrorString(e)); rc = DGLLB_ERR_CUDA; break; }
        if (n_seeds > 0 && out_col) {
            const long long blocks = (n_seeds * 32 + tb - 1) / tb;
            sample_fill_kernel<<<static_cast<unsigned>(blocks), tb, 0, st>>>(
                row_ptr, row_ptr_is64, col_idx, seeds, seeds_is64, n_seeds, fanout, rng_seed, rng_offset, out_row_ptr,
                out_col);
            g_launch_count.fetch_add(1);
        }
        e = cudaGetLastError();
        if (e != cudaSuccess) { set_error("sample_neighbors: %s", cudaGetErrorString(e)); rc = DGLLB_ERR_CUDA; }
    } while (0);
    cudaFreeAsync(ws, st);
    return rc;
}

extern "C" int dgllb_sample_neighbors(const void* row_ptr, int row_ptr_is64, const int32_t* col_idx,
                                      const void* seeds, int seeds_is64, int64_t n_seeds, int fanout,
                                      uint64_t rng_seed, int32_t* out_row_ptr, int32_t* out_col,
                                      void* stream) {
    return sample_neighbors_impl(row_ptr, row_ptr_is64, col_idx, seeds, seeds_is64, n_seeds, fanout, rng_seed, nullptr,
                                 out_row_ptr, out_col, stream);
}

extern "C" int dgllb_sample_neighbors_cap(const void* row_ptr, int row_ptr_is64, const int32_t* col_idx,
                                          const void* seeds, int seeds_is64, int64_t n_seeds_cap, int fanout,
                                          uint64_t rng_seed, const uint64_t* rng_offset, int32_t* out_row_ptr,
                                          int32_t* out_col, void* stream) {
    return sample_neighbors_impl(row_ptr, row_ptr_is64, col_idx, seeds, seeds_is64, n_seeds_cap, fanout, rng_seed,
                                 reinterpret_cast<const unsigned long long*>(rng_offset), out_row_ptr, out_col, stream);
}
