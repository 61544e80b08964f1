// block.cu — device-side block (message-flow-graph) construction: dst-first compaction of a sampled edge list.
//
// Counterpart of the relabelling DGL's to_block performs for the reference's GPU-Accelerator scripts (blocks whose src
// id space starts with the dst nodes in order, GPU Accelerator/MQGCN.py:45,48) and of sugbraph's unique-node
// bookkeeping (dgll/sampling/base_sampler.py:81).  Given the dst ids (unique) and the sampled neighbours' GLOBAL ids
// (CSR by destination from dgllb_sample_neighbors), it produces
//     src_ids[num_src]   unique ids of cat(dst_ids, neighbours) in FIRST-OCCURRENCE order (so src_ids[:n_dst] == dst_ids)
//     col_local[nnz]     the neighbours relabelled into that space
// deterministically and without sorting: an open-addressing hash table records, per id, the smallest position at which
// it occurs (atomicMin); positions that are first occurrences are flagged, an exclusive scan of the flags is the new
// label, and every neighbour looks its label up through the table.  One device->host read-back (num_src, nnz) instead of
// the sort-based unique + three read-backs of the torch formulation (profiles: 1.0 ms -> see DESIGN.md §9 of host time).
#include "common.cuh"
#include <cub/device/device_scan.cuh>

namespace dgllb {

__device__ __forceinline__ unsigned hash64(long long k, unsigned mask) {
    unsigned long long z = static_cast<unsigned long long>(k) * 0x9E3779B97F4A7C15ull;
    return static_cast<unsigned>(z >> 32) & mask;
}

__device__ __forceinline__ long long block_id_at(const long long* dst_ids, const int* nbr, long long n_dst, long long pos) {
    return pos < n_dst ? dst_ids[pos] : static_cast<long long>(nbr[pos - n_dst]);
}

// total = n_dst + nnz where nnz = row_ptr[n_dst] is read on the device
__global__ void block_insert_kernel(const long long* __restrict__ dst_ids, const int* __restrict__ nbr,
                                    const int* __restrict__ row_ptr, long long n_dst, long long* keys, int* minpos,
                                    unsigned mask) {
    const long long total = n_dst + row_ptr[n_dst];
    for (long long pos = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; pos < total;
         pos += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long id = block_id_at(dst_ids, nbr, n_dst, pos);
        if (id < 0) continue;  // padding slot of a fixed-capacity dst array
        unsigned slot = hash64(id, mask);
        while (true) {
            const long long prev = atomicCAS(reinterpret_cast<unsigned long long*>(keys + slot),
                                             static_cast<unsigned long long>(-1ll), static_cast<unsigned long long>(id));
            if (prev == -1ll || prev == id) break;
            slot = (slot + 1) & mask;
        }
        atomicMin(minpos + slot, static_cast<int>(pos));
    }
}

__device__ __forceinline__ int block_lookup(const long long* keys, const int* minpos, unsigned mask, long long id) {
    unsigned slot = hash64(id, mask);
    while (keys[slot] != id) slot = (slot + 1) & mask;
    return minpos[slot];
}

__global__ void block_flag_kernel(const long long* __restrict__ dst_ids, const int* __restrict__ nbr,
                                  const int* __restrict__ row_ptr, long long n_dst, long long cap,
                                  const long long* __restrict__ keys, const int* __restrict__ minpos, unsigned mask,
                                  int* __restrict__ flags, int* __restrict__ first_of) {
    const long long total = n_dst + row_ptr[n_dst];
    for (long long pos = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; pos < cap;
         pos += static_cast<long long>(gridDim.x) * blockDim.x) {
        int f = 0, mp = 0;
        if (pos < total) {
            const long long id = block_id_at(dst_ids, nbr, n_dst, pos);
            if (id >= 0) {
                mp = block_lookup(keys, minpos, mask, id);
                f = mp == static_cast<int>(pos);
            }
        }
        flags[pos] = f;
        first_of[pos] = mp;
    }
}

__global__ void block_emit_kernel(const long long* __restrict__ dst_ids, const int* __restrict__ nbr,
                                  const int* __restrict__ row_ptr, long long n_dst, long long cap,
                                  const int* __restrict__ flags, const int* __restrict__ rank,
                                  const int* __restrict__ first_of, long long* __restrict__ src_ids,
                                  int* __restrict__ col_local, int* __restrict__ counts_out, int col_pad,
                                  int n_counts) {
    const long long nnz = row_ptr[n_dst];
    const long long total = n_dst + nnz;
    const long long limit = col_pad >= 0 ? cap : total;   // capacity mode also fills the unused edge slots
    if (total == 0 && blockIdx.x == 0 && threadIdx.x == 0)
        for (int k = 0; k < n_counts; ++k) counts_out[k] = 0;
    for (long long pos = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; pos < limit;
         pos += static_cast<long long>(gridDim.x) * blockDim.x) {
        if (pos >= total) {
            col_local[pos - n_dst] = col_pad;
            continue;
        }
        if (flags[pos]) src_ids[rank[pos]] = block_id_at(dst_ids, nbr, n_dst, pos);
        if (pos >= n_dst) col_local[pos - n_dst] = rank[first_of[pos]];
        if (pos == total - 1) {
            counts_out[0] = rank[pos] + flags[pos];  // num_src
            counts_out[1] = static_cast<int>(nnz);
            // valid (non-negative) dst ids come first and are unique: each is its own first occurrence
            if (n_counts > 2) counts_out[2] = n_dst < total ? rank[n_dst] : rank[pos] + flags[pos];
        }
    }
}

}  // namespace dgllb

using namespace dgllb;

static int build_block_impl(const int64_t* dst_ids, int64_t n_dst, const int32_t* row_ptr, const int32_t* nbr_global,
                            int64_t nnz_cap, int64_t* src_ids, int32_t* col_local, int32_t* counts_out, int col_pad,
                            int n_counts, void* stream) {
    DGLLB_REQUIRE(n_dst >= 0 && nnz_cap >= 0, "build_block: negative size");
    DGLLB_REQUIRE(n_dst + nnz_cap < (1ll << 30), "build_block: block too large for 32-bit positions");
    DGLLB_REQUIRE(row_ptr && src_ids && counts_out && (n_dst == 0 || dst_ids) && (nnz_cap == 0 || (nbr_global && col_local)),
                  "build_block: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    DevInfo di;
    int rc = get_devinfo(&di);
    if (rc != DGLLB_OK) return rc;
    const long long cap = n_dst + nnz_cap;
    if (cap == 0) {
        DGLLB_CUDA_TRY(cudaMemsetAsync(counts_out, 0, n_counts * sizeof(int), st));
        return DGLLB_OK;
    }
    unsigned tsize = 1024;
    while (tsize < 2 * cap) tsize <<= 1;
    const unsigned mask = tsize - 1;
    size_t cub_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, static_cast<int*>(nullptr), static_cast<int*>(nullptr),
                                  static_cast<int>(cap), st);
    auto al = [](size_t b) { return (b + 255) & ~static_cast<size_t>(255); };
    const size_t b_keys = al(sizeof(long long) * tsize), b_min = al(sizeof(int) * tsize), b_i = al(sizeof(int) * cap);
    char* ws = nullptr;
    DGLLB_CUDA_TRY(cudaMallocAsync(&ws, b_keys + b_min + 3 * b_i + cub_bytes, st));
    long long* keys = reinterpret_cast<long long*>(ws);
    int* minpos = reinterpret_cast<int*>(ws + b_keys);
    int* flags = reinterpret_cast<int*>(ws + b_keys + b_min);
    int* rank = reinterpret_cast<int*>(ws + b_keys + b_min + b_i);
    int* first_of = reinterpret_cast<int*>(ws + b_keys + b_min + 2 * b_i);
    void* cub_ws = ws + b_keys + b_min + 3 * b_i;
    do {
        cudaError_t e = cudaMemsetAsync(keys, 0xFF, sizeof(long long) * tsize, st);      // all keys = -1
        if (e == cudaSuccess) e = cudaMemsetAsync(minpos, 0x7F, sizeof(int) * tsize, st); // large positive
        if (e == cudaSuccess && n_counts > 2)   // capacity mode: unused src slots read as padding (-1) downstream
            e = cudaMemsetAsync(src_ids, 0xFF, sizeof(long long) * cap, st);
        if (e != cudaSuccess) { set_error("build_block: %s", cudaGetErrorString(e)); rc = DGLLB_ERR_CUDA; break; }
        const int tb = 256;
        long long blocks = (cap + tb - 1) / tb;
        if (blocks > static_cast<long long>(di.sm_count) * 16) blocks = static_cast<long long>(di.sm_count) * 16;
        const unsigned g = static_cast<unsigned>(blocks);
        block_insert_kernel<<<g, tb, 0, st>>>(reinterpret_cast<const long long*>(dst_ids), nbr_global, row_ptr, n_dst, keys,
                                              minpos, mask);
        block_flag_kernel<<<g, tb, 0, st>>>(reinterpret_cast<const long long*>(dst_ids), nbr_global, row_ptr, n_dst, cap,
                                            keys, minpos, mask, flags, first_of);
        e = cub::DeviceScan::ExclusiveSum(cub_ws, cub_bytes, flags, rank, static_cast<int>(cap), st);
        if (e != cudaSuccess) { set_error("build_block: scan: %s", cudaGetErrorString(e)); rc = DGLLB_ERR_CUDA; break; }
        block_emit_kernel<<<g, tb, 0, st>>>(reinterpret_cast<const long long*>(dst_ids), nbr_global, row_ptr, n_dst, cap,
                                            flags, rank, first_of, reinterpret_cast<long long*>(src_ids), col_local,
                                            counts_out, col_pad, n_counts);
        g_launch_count.fetch_add(4);
        e = cudaGetLastError();
        if (e != cudaSuccess) { set_error("build_block: %s", cudaGetErrorString(e)); rc = DGLLB_ERR_CUDA; }
    } while (0);
    cudaFreeAsync(ws, st);
    return rc;
}

extern "C" int dgllb_build_block(const int64_t* dst_ids, int64_t n_dst, const int32_t* row_ptr,
                                 const int32_t* nbr_global, int64_t nnz_cap, int64_t* src_ids, int32_t* col_local,
                                 int32_t* counts_out, void* stream) {
    return build_block_impl(dst_ids, n_dst, row_ptr, nbr_global, nnz_cap, src_ids, col_local, counts_out, -1, 2, stream);
}

extern "C" int dgllb_build_block_cap(const int64_t* dst_ids, int64_t n_dst_cap, const int32_t* row_ptr,
                                     const int32_t* nbr_global, int64_t nnz_cap, int col_pad, int64_t* src_ids,
                                     int32_t* col_local, int32_t* counts_out, void* stream) {
    DGLLB_REQUIRE(col_pad >= 0, "build_block_cap: col_pad must be >= 0 (the padding column id of the consumer)");
    return build_block_impl(dst_ids, n_dst_cap, row_ptr, nbr_global, nnz_cap, src_ids, col_local, counts_out, col_pad, 3,
                            stream);
}
