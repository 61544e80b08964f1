// gather.cu — sampled-block feature gather for sm_100a.
//
// Replaces: features[nodes] (dgll/data/dgraph.py:105), the HBM-cache gather and
// masked scatter of GraphCacheServer.fetch_data / fetch_from_cache
// (dgll/FeatureCache/storage.py:151-210).
//
// Design: a row gather is a pure byte move, so rows travel as TMA bulk copies
// (cp.async.bulk global->shared with an mbarrier transaction count, then
// cp.async.bulk shared->global) — no registers, one elected lane per warp
// drives a ring of kSlots shared-memory slots, so every warp keeps kSlots row
// pieces in flight.  SASS: UBLKCP.  Rows whose base/stride are not 16-byte
// multiples fall back to a vector/scalar LDG copy.
// Algorithmic bytes: M * (id_bytes + 2 * row_bytes).
#include "common.cuh"

namespace dgllb {

constexpr int kGatherWarps = 8;      // warps per CTA
constexpr int kMaxSlots = 16;        // ring depth per warp (upper bound)
constexpr int kSlotBytes = 4096;     // piece size (rows longer than this are split)
constexpr int kGatherSmem = 200 * 1024;

struct GatherSrc {
    // resolves the source address of output row i
    const char* table;
    long long stride;
    const void* ids;
    int ids64;
    // optional cache split (GraphCacheServer): when gpu_flag != nullptr
    const unsigned char* gpu_flag;
    const long long* local2cache;
    const long long* nid_map;
    const char* host_table;
    long long host_stride;
};

__device__ __forceinline__ const char* resolve_row(const GatherSrc& s, long long i, bool* is_host) {
    const long long id = s.ids64 ? reinterpret_cast<const long long*>(s.ids)[i]
                                 : static_cast<long long>(reinterpret_cast<const int*>(s.ids)[i]);
    *is_host = false;
    if (s.gpu_flag) {
        if (s.gpu_flag[id]) return s.table + s.local2cache[id] * s.stride;
        *is_host = true;
        const long long hid = s.nid_map ? s.nid_map[id] : id;
        return s.host_table + hid * s.host_stride;
    }
    return s.table + id * s.stride;
}

// TMA bulk-copy gather.  grid = persistent CTAs; every warp owns a CONTIGUOUS range of (row, piece) items so
// its ids are read coalesced, 32 at a time, by all lanes (the first version let the single issuing lane load
// each id itself and stalled ~1 us per row on that dependent load — profiles/r01_bin_gather.txt).  The lanes
// park (src, dst, bytes) of the next 64 items in shared memory; lane 0 runs the copy pipeline: a ring of
// `slots` shared-memory slots, loads `slots` items ahead of the stores.
__global__ void __launch_bounds__(kGatherWarps * 32)
gather_bulk_kernel(const GatherSrc src, char* __restrict__ out, long long out_stride, long long n_rows,
                   long long row_bytes, int pieces_per_row, int slots, int slot_bytes, int piece_bytes) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bars[kGatherWarps][kMaxSlots];
    __shared__ unsigned long long m_src[kGatherWarps][64], m_dst[kGatherWarps][64];
    __shared__ uint32_t m_bytes[kGatherWarps][64];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* ring = smem + static_cast<size_t>(warp) * slots * slot_bytes;
    if (lane == 0) {
        for (int s = 0; s < slots; ++s) mbar_init(&bars[warp][s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    const long long n_items = n_rows * pieces_per_row;
    const long long n_warps = static_cast<long long>(gridDim.x) * kGatherWarps;
    const long long per_warp = (n_items + n_warps - 1) / n_warps;
    const long long gw = static_cast<long long>(blockIdx.x) * kGatherWarps + warp;
    const long long beg = gw * per_warp;
    if (beg >= n_items) return;
    const long long count = min(per_warp, n_items - beg);

    auto prepare = [&](long long batch) {
        const long long n = batch * 32 + lane;
        if (n < count) {
            const long long item = beg + n;
            const long long row = item / pieces_per_row;
            const int piece = static_cast<int>(item - row * pieces_per_row);
            const long long off = static_cast<long long>(piece) * piece_bytes;
            bool is_host;
            const char* g = resolve_row(src, row, &is_host) + off;
            const int k = static_cast<int>(n & 63);
            m_src[warp][k] = reinterpret_cast<unsigned long long>(g);
            m_dst[warp][k] = reinterpret_cast<unsigned long long>(out + row * out_stride + off);
            m_bytes[warp][k] = static_cast<uint32_t>(min(static_cast<long long>(piece_bytes), row_bytes - off));
        }
    };
    auto issue_load = [&](long long n) {
        const int k = static_cast<int>(n & 63), slot = static_cast<int>(n % slots);
        const uint32_t bytes = m_bytes[warp][k];
        mbar_expect_tx(&bars[warp][slot], bytes);
        bulk_g2s(ring + slot * slot_bytes, reinterpret_cast<const void*>(m_src[warp][k]), bytes, &bars[warp][slot]);
    };
    auto issue_store = [&](long long n) {
        const int k = static_cast<int>(n & 63), slot = static_cast<int>(n % slots);
        bulk_s2g(reinterpret_cast<void*>(m_dst[warp][k]), ring + slot * slot_bytes, m_bytes[warp][k]);
    };

    prepare(0);
    prepare(1);
    __syncwarp();
    if (lane == 0) {
        const long long pre = min(static_cast<long long>(slots), count);
        for (long long n = 0; n < pre; ++n) issue_load(n);
    }
    const long long n_batches = (count + 31) / 32;
    for (long long b = 0; b < n_batches; ++b) {
        if (lane == 0) {
            const long long hi = min(count, b * 32 + 32);
            for (long long n = b * 32; n < hi; ++n) {
                const int slot = static_cast<int>(n % slots);
                mbar_wait(&bars[warp][slot], static_cast<uint32_t>((n / slots) & 1));
                issue_store(n);
                // refill the slot of the PREVIOUS store once that store has drained its smem
                if (n >= 1 && (n - 1) + slots < count) {
                    bulk_wait_read<1>();
                    issue_load((n - 1) + slots);
                }
            }
        }
        __syncwarp();
        prepare(b + 2);  // reuses the metadata slots of batch b
        __syncwarp();
    }
    if (lane == 0) bulk_wait_all();
}

// generic fallback: one warp per row, 16-byte / 4-byte / 1-byte moves by alignment
__global__ void __launch_bounds__(256)
gather_ldg_kernel(const GatherSrc src, char* __restrict__ out, long long out_stride, long long n_rows,
                  long long row_bytes, long long* miss_count) {
    const int lane = threadIdx.x & 31;
    const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (row >= n_rows) return;
    bool is_host;
    const char* s = resolve_row(src, row, &is_host);
    char* d = out + row * out_stride;
    if (miss_count && is_host && lane == 0) atomicAdd(reinterpret_cast<unsigned long long*>(miss_count), 1ull);
    const uintptr_t al = reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d) |
                         static_cast<uintptr_t>(row_bytes);
    if ((al & 15) == 0) {
        const long long n16 = row_bytes >> 4;
        for (long long i = lane; i < n16; i += 32)
            reinterpret_cast<uint4*>(d)[i] = ldg_nc_u4(reinterpret_cast<const uint4*>(s) + i);
    } else if ((al & 3) == 0) {
        const long long n4 = row_bytes >> 2;
        for (long long i = lane; i < n4; i += 32)
            reinterpret_cast<uint32_t*>(d)[i] = __ldg(reinterpret_cast<const uint32_t*>(s) + i);
    } else {
        for (long long i = lane; i < row_bytes; i += 32) d[i] = s[i];
    }
}

}  // namespace dgllb

using namespace dgllb;

static int launch_gather(const GatherSrc& src, void* out, int64_t out_stride, int64_t n_rows,
                         int64_t row_bytes, bool allow_bulk, long long* miss_count, cudaStream_t st) {
    const bool bulk_ok = allow_bulk && aligned16(src.table) && aligned16(out) && (src.stride % 16 == 0) &&
                         (out_stride % 16 == 0) && (row_bytes % 16 == 0) && row_bytes >= 16;
    if (bulk_ok) {
        DevInfo di;
        int rc = get_devinfo(&di);
        if (rc != DGLLB_OK) return rc;
        const int piece_bytes = static_cast<int>(row_bytes < kSlotBytes ? row_bytes : kSlotBytes);
        const int pieces = static_cast<int>((row_bytes + piece_bytes - 1) / piece_bytes);
        const int slot_bytes = (piece_bytes + 127) & ~127;
        int slots = kGatherSmem / (kGatherWarps * slot_bytes);
        if (slots > kMaxSlots) slots = kMaxSlots;
        const size_t smem = static_cast<size_t>(kGatherWarps) * slots * slot_bytes;
        // opt in to >48 KB dynamic shared memory (idempotent)
        DGLLB_CUDA_TRY(cudaFuncSetAttribute(gather_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            kGatherSmem));
        const long long items = n_rows * pieces;
        long long blocks = (items + kGatherWarps * 8 - 1) / (kGatherWarps * 8);  // >= 8 items per warp
        const long long max_blocks = static_cast<long long>(di.sm_count);         // ~200 KB smem => 1 CTA/SM
        if (blocks > max_blocks) blocks = max_blocks;
        if (blocks < 1) blocks = 1;
        gather_bulk_kernel<<<static_cast<unsigned>(blocks), kGatherWarps * 32, smem, st>>>(
            src, static_cast<char*>(out), out_stride, n_rows, row_bytes, pieces, slots, slot_bytes, piece_bytes);
        DGLLB_LAUNCH_CHECK();
        return DGLLB_OK;
    }
    const long long blocks = (n_rows * 32 + 255) / 256;
    DGLLB_REQUIRE(blocks < (1ll << 31), "gather: grid too large");
    gather_ldg_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(src, static_cast<char*>(out), out_stride,
                                                                     n_rows, row_bytes, miss_count);
    DGLLB_LAUNCH_CHECK();
    return DGLLB_OK;
}

extern "C" int dgllb_gather_rows(const void* table, int64_t table_stride_bytes, const void* ids,
                                 int ids_is64, void* out, int64_t out_stride_bytes, int64_t n_rows,
                                 int64_t row_bytes, void* stream) {
    if (n_rows == 0 || row_bytes == 0) return DGLLB_OK;
    DGLLB_REQUIRE(table && ids && out, "gather_rows: null pointer");
    DGLLB_REQUIRE(n_rows > 0 && row_bytes > 0, "gather_rows: negative size");
    DGLLB_REQUIRE(table_stride_bytes >= row_bytes && out_stride_bytes >= row_bytes,
                  "gather_rows: stride smaller than row_bytes");
    GatherSrc s;
    s.table = static_cast<const char*>(table);
    s.stride = table_stride_bytes;
    s.ids = ids;
    s.ids64 = ids_is64;
    s.gpu_flag = nullptr;
    s.local2cache = nullptr;
    s.nid_map = nullptr;
    s.host_table = nullptr;
    s.host_stride = 0;
    return launch_gather(s, out, out_stride_bytes, n_rows, row_bytes, true, nullptr,
                         static_cast<cudaStream_t>(stream));
}

extern "C" int dgllb_gather_rows_cached(const void* cache_table, int64_t cache_stride_bytes,
                                        const void* host_table, int64_t host_stride_bytes,
                                        const int64_t* ids, const uint8_t* gpu_flag,
                                        const int64_t* localid2cacheid, const int64_t* nid_map,
                                        void* out, int64_t out_stride_bytes, int64_t n_rows,
                                        int64_t row_bytes, int64_t* miss_count, void* stream) {
    if (n_rows == 0 || row_bytes == 0) return DGLLB_OK;
    DGLLB_REQUIRE(ids && gpu_flag && localid2cacheid && out, "gather_rows_cached: null pointer");
    DGLLB_REQUIRE(cache_table || host_table, "gather_rows_cached: no table");
    GatherSrc s;
    s.table = static_cast<const char*>(cache_table);
    s.stride = cache_stride_bytes;
    s.ids = ids;
    s.ids64 = 1;
    s.gpu_flag = gpu_flag;
    s.local2cache = reinterpret_cast<const long long*>(localid2cacheid);
    s.nid_map = reinterpret_cast<const long long*>(nid_map);
    s.host_table = static_cast<const char*>(host_table);
    s.host_stride = host_stride_bytes;
    // misses read pinned host memory through the PCIe aperture: plain LDGs, not TMA
    return launch_gather(s, out, out_stride_bytes, n_rows, row_bytes, false,
                         reinterpret_cast<long long*>(miss_count), static_cast<cudaStream_t>(stream));
}
