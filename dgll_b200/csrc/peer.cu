// peer.cu — feature tables sharded across the GPUs of one NVSwitch box, read in place over NVLink.
//
// The halo exchange of a node-range-partitioned feature table (SURVEY.md §8 e) does not need a collective on a
// B200 box: every GPU can map every peer's shard (CUDA IPC) and the gather kernel issues the remote 128-bit loads
// itself — NVSwitch gives each GPU full bandwidth to each peer.  One kernel launch replaces bucket-by-owner +
// all_to_all(counts) + all_to_all(ids) + owner-side gather + all_to_all(rows) + un-permute, and nothing is read
// back to the host.  row id -> shard = id / rows_per_shard (node-range partition, no lookup table).
//   dgllb_ipc_export / dgllb_ipc_import / dgllb_ipc_release : share one device allocation between processes
//   dgllb_gather_rows_sharded                              : out[i] = shard[id/part][id%part]  (exact byte copy)
// Evidence of the peer path: LDG.E.128 on the mapped peer pointers inside gather_sharded_kernel.
#include "common.cuh"
#include <cuda.h>
#include <string.h>

namespace dgllb {

__global__ void __launch_bounds__(256)
gather_sharded_kernel(const char* const* __restrict__ shard_ptrs, long long rows_per_shard, long long stride,
                      const void* __restrict__ ids, int ids64, char* __restrict__ out, long long out_stride,
                      long long n_rows, long long row_bytes) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const long long n16 = row_bytes >> 4;
    for (long long row = warp0; row < n_rows; row += n_warps) {
        const long long id = ids64 ? reinterpret_cast<const long long*>(ids)[row]
                                   : static_cast<long long>(reinterpret_cast<const int*>(ids)[row]);
        char* __restrict__ d = out + row * out_stride;
        if (id < 0) {   // padding slot of a fixed-capacity id array: no remote read, the row reads as zeros
            for (long long i = lane; i < row_bytes; i += 32) d[i] = 0;
            continue;
        }
        const long long shard = id / rows_per_shard;
        const char* __restrict__ s = shard_ptrs[shard] + (id - shard * rows_per_shard) * stride;
        if (((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d) | static_cast<uintptr_t>(row_bytes)) & 15) == 0) {
            // up to 4 x 16 B per lane in flight (rows up to 2 KB move in one round)
            for (long long i = lane; i < n16; i += 128) {
                uint4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (i + 32 * u < n16) v[u] = ldg_nc_u4(reinterpret_cast<const uint4*>(s) + i + 32 * u);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (i + 32 * u < n16) reinterpret_cast<uint4*>(d)[i + 32 * u] = v[u];
            }
        } else {
            for (long long i = lane; i < row_bytes; i += 32) d[i] = s[i];
        }
    }
}

typedef CUresult (*GetRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
static GetRangeFn get_range_fn() {
    static GetRangeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<GetRangeFn>(p);
    }
    return fn;
}

}  // namespace dgllb

using namespace dgllb;

extern "C" int dgllb_gather_rows_sharded(const void* const* shard_ptrs, int n_shards, int64_t rows_per_shard,
                                         int64_t stride_bytes, const void* ids, int ids_is64, void* out,
                                         int64_t out_stride_bytes, int64_t n_rows, int64_t row_bytes, void* stream) {
    if (n_rows == 0 || row_bytes == 0) return DGLLB_OK;
    DGLLB_REQUIRE(shard_ptrs && ids && out, "gather_rows_sharded: null pointer");
    DGLLB_REQUIRE(n_shards >= 1 && rows_per_shard >= 1, "gather_rows_sharded: bad shard geometry");
    DGLLB_REQUIRE(n_rows > 0 && row_bytes > 0 && stride_bytes >= row_bytes && out_stride_bytes >= row_bytes,
                  "gather_rows_sharded: bad sizes");
    DevInfo di;
    int rc = get_devinfo(&di);
    if (rc != DGLLB_OK) return rc;
    long long blocks = (n_rows * 32 + 255) / 256;
    const long long cap = static_cast<long long>(di.sm_count) * 8 * 4;   // persistent-ish: 4 waves of 8 CTAs/SM
    if (blocks > cap) blocks = cap;
    gather_sharded_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const char* const*>(shard_ptrs), rows_per_shard, stride_bytes, ids, ids_is64,
        static_cast<char*>(out), out_stride_bytes, n_rows, row_bytes);
    DGLLB_LAUNCH_CHECK();
    return DGLLB_OK;
}

extern "C" int dgllb_ipc_export(const void* dev_ptr, unsigned char* handle64, int64_t* offset) {
    DGLLB_REQUIRE(dev_ptr && handle64 && offset, "ipc_export: null pointer");
    GetRangeFn range = get_range_fn();
    DGLLB_REQUIRE(range != nullptr, "ipc_export: cuMemGetAddressRange unavailable");
    CUdeviceptr base = 0;
    size_t size = 0;
    CUresult r = range(&base, &size, reinterpret_cast<CUdeviceptr>(dev_ptr));
    if (r != CUDA_SUCCESS) {
        set_error("ipc_export: cuMemGetAddressRange failed (%d)", static_cast<int>(r));
        return DGLLB_ERR_CUDA;
    }
    cudaIpcMemHandle_t h;
    DGLLB_CUDA_TRY(cudaIpcGetMemHandle(&h, reinterpret_cast<void*>(base)));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handle64, &h, 64);
    *offset = static_cast<int64_t>(reinterpret_cast<CUdeviceptr>(dev_ptr) - base);
    return DGLLB_OK;
}

extern "C" int dgllb_ipc_import(const unsigned char* handle64, int64_t offset, void** dev_ptr_out) {
    DGLLB_REQUIRE(handle64 && dev_ptr_out, "ipc_import: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* base = nullptr;
    DGLLB_CUDA_TRY(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    *dev_ptr_out = static_cast<char*>(base) + offset;
    return DGLLB_OK;
}

extern "C" int dgllb_ipc_release(void* dev_ptr, int64_t offset) {
    if (!dev_ptr) return DGLLB_OK;
    DGLLB_CUDA_TRY(cudaIpcCloseMemHandle(static_cast<char*>(dev_ptr) - offset));
    return DGLLB_OK;
}
