// common.cuh — shared host/device helpers for the dgll_b200 C-ABI library.
// sm_100a only (B200).  No torch types anywhere in csrc/.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/dgll_b200.h"

namespace dgllb {

// ---------------------------------------------------------------- errors --
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launch_count;

struct DevInfo {
    int sm_count;
    int cc_major, cc_minor;
    long long l2_bytes;
    int max_smem_optin;
};
// Cached properties of the current device (thread-safe, per device index).
int get_devinfo(DevInfo* out);

// Tuning options: integers initialised ONCE from the environment (DGLLB_<NAME>) and changed afterwards only through
// dgllb_set_option — no getenv() on any launch path.  0 always means "library default".
enum Opt {
    OPT_SPMM_KERNEL = 0,   // 0 auto, 1 rowsplit, 2 stream, 3 wholerow
    OPT_SPMM_TB,           // rowsplit kernel block size
    OPT_ROWS_TB, OPT_ROWS_NS, OPT_ROWS_D,   // whole-row kernel: block size, slabs per warp, window depth
    OPT_ROWS_STREAM,       // whole-row kernel across rows: 0 auto, 1 off, n >= 2 = n rows per warp
    OPT_ROWS_SHARDED_BPS,  // sharded aggregation: resident 64-thread blocks per SM (5..16; 0 = no limit)
    OPT_GAT_KERNEL,        // 0 auto, 1 generic (lane-group), 2 whole-row
    OPT_GAT_ROW_WARPS, OPT_GAT_BWD_TB,
    OPT_GAT_BWD_KERNEL,    // 0 auto, 1 two-pass (CSR then CSR^T), 2 fused single pass over CSR^T
    OPT_GAT_BWD_DEPTH,     // fused backward: gradient rows in flight per lane (4 or 8; 0 = default)
    OPT_BIN_TB,
    OPT_GEMM_KERNEL,       // 0 auto; 3 = TF32 one-tile-per-CTA kernel, 4 = TF32 persistent kernel, 5 = precision 0 always SIMT FMA
    OPT_NVTX,              // 1 = emit NVTX ranges around the entry points that mirror the reference's ranges
    OPT_COUNT
};
int opt_get(Opt o);

#define DGLLB_CUDA_TRY(expr)                                                        \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) {                                                    \
            dgllb::set_error("%s failed at %s:%d: %s", #expr, __FILE__, __LINE__,   \
                             cudaGetErrorString(_e));                               \
            return DGLLB_ERR_CUDA;                                                  \
        }                                                                           \
    } while (0)

#define DGLLB_REQUIRE(cond, ...)                 \
    do {                                         \
        if (!(cond)) {                           \
            dgllb::set_error(__VA_ARGS__);       \
            return DGLLB_ERR_INVALID;            \
        }                                        \
    } while (0)

// call after every kernel launch
#define DGLLB_LAUNCH_CHECK()                                                        \
    do {                                                                            \
        dgllb::g_launch_count.fetch_add(1, std::memory_order_relaxed);              \
        cudaError_t _e = cudaGetLastError();                                        \
        if (_e != cudaSuccess) {                                                    \
            dgllb::set_error("kernel launch failed at %s:%d: %s", __FILE__,         \
                             __LINE__, cudaGetErrorString(_e));                     \
            return DGLLB_ERR_CUDA;                                                  \
        }                                                                           \
    } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ------------------------------------------------------ device helpers ----
#ifdef __CUDACC__

// 128-bit streaming load through the read-only path, no L1 allocation: feature
// rows are reused at L2 (several destinations share a source) but not in L1.
__device__ __forceinline__ float4 ldg_nc_f4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ldg_nc_u4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ float ldg_nc_f1(const float* p) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_cs_f4(float* p, float4 v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}

__device__ __forceinline__ float bf16lo_to_f32(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi_to_f32(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

template <typename RP>
__device__ __forceinline__ long long ld_rowptr(const RP* p, long long i) {
    return static_cast<long long>(p[i]);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- mbarrier + bulk-async-copy (TMA engine, SASS UBLKCP) helpers ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

#endif  // __CUDACC__

}  // namespace dgllb
