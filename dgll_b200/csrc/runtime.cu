// runtime.cu — error string, device-info cache, launch counter.
#include "common.cuh"
#include <mutex>
#include <string.h>
#include <stdlib.h>

namespace dgllb {

static thread_local char t_err[512] = "";
std::atomic<long long> g_launch_count{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

static const int kMaxDev = 64;
static DevInfo g_info[kMaxDev];
static std::atomic<int> g_info_ok[kMaxDev];
static std::mutex g_info_mu;

int get_devinfo(DevInfo* out) {
    int dev = 0;
    DGLLB_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDev) {
        set_error("device index %d out of range", dev);
        return DGLLB_ERR_INVALID;
    }
    if (!g_info_ok[dev].load(std::memory_order_acquire)) {
        std::lock_guard<std::mutex> lk(g_info_mu);
        if (!g_info_ok[dev].load(std::memory_order_relaxed)) {
            DevInfo d;
            int v = 0;
            DGLLB_CUDA_TRY(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev));
            DGLLB_CUDA_TRY(cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
            DGLLB_CUDA_TRY(cudaDeviceGetAttribute(&d.cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
            DGLLB_CUDA_TRY(cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, dev));
            d.l2_bytes = v;
            DGLLB_CUDA_TRY(cudaDeviceGetAttribute(&d.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
            g_info[dev] = d;
            // Workspace comes from the stream-ordered allocator.  Its default release threshold is 0: every
            // synchronisation hands the pool back to the OS and the next cudaMallocAsync pays milliseconds
            // (measured: 4.6 ms per dgllb_csr_transpose inside a training step).  Keep freed blocks cached.
            cudaMemPool_t pool;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
                unsigned long long keep = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            }
            g_info[dev].sm_count = d.sm_count;
            g_info_ok[dev].store(1, std::memory_order_release);
        }
    }
    *out = g_info[dev];
    return DGLLB_OK;
}

// ---------------------------------------------------------------- options --
struct OptDesc { const char* name; const char* env; const char* const* words; };
static const char* const kSpmmWords[] = {"auto", "rowsplit", "stream", "wholerow", nullptr};
static const char* const kGatWords[] = {"auto", "group", "row", nullptr};
static const char* const kGatBwdWords[] = {"auto", "twopass", "fused", nullptr};
static const OptDesc kOpts[OPT_COUNT] = {
    {"spmm_kernel", "DGLLB_SPMM_KERNEL", kSpmmWords},
    {"spmm_tb", "DGLLB_SPMM_TB", nullptr},
    {"rows_tb", "DGLLB_ROWS_TB", nullptr},
    {"rows_ns", "DGLLB_ROWS_NS", nullptr},
    {"rows_d", "DGLLB_ROWS_D", nullptr},
    {"rows_stream", "DGLLB_ROWS_STREAM", nullptr},
    {"rows_sharded_bps", "DGLLB_ROWS_SHARDED_BPS", nullptr},
    {"gat_kernel", "DGLLB_GAT_KERNEL", kGatWords},
    {"gat_row_warps", "DGLLB_GAT_ROW_WARPS", nullptr},
    {"gat_bwd_tb", "DGLLB_GAT_BWD_TB", nullptr},
    {"gat_bwd_kernel", "DGLLB_GAT_BWD_KERNEL", kGatBwdWords},
    {"gat_bwd_depth", "DGLLB_GAT_BWD_DEPTH", nullptr},
    {"bin_tb", "DGLLB_BIN_TB", nullptr},
    {"gemm_kernel", "DGLLB_GEMM_KERNEL", nullptr},
    {"nvtx", "DGLLB_NVTX", nullptr},
};
static std::atomic<int> g_opt[OPT_COUNT];
static std::once_flag g_opt_once;

static bool opt_parse(const OptDesc& d, const char* v, int* out) {
    if (!v || !*v) { *out = 0; return true; }
    if (d.words)
        for (int k = 0; d.words[k]; ++k)
            if (strcmp(d.words[k], v) == 0) { *out = k; return true; }
    char* end = nullptr;
    const long x = strtol(v, &end, 10);
    if (end == v || *end) return false;
    *out = static_cast<int>(x);
    return true;
}

static void opt_init() {
    for (int k = 0; k < OPT_COUNT; ++k) {
        int v = 0;
        if (const char* e = getenv(kOpts[k].env)) opt_parse(kOpts[k], e, &v);
        g_opt[k].store(v, std::memory_order_relaxed);
    }
}

int opt_get(Opt o) {
    std::call_once(g_opt_once, opt_init);
    return g_opt[o].load(std::memory_order_relaxed);
}

}  // namespace dgllb

extern "C" {

int dgllb_version(void) { return 2000; }

int dgllb_set_option(const char* name, const char* value) {
    using namespace dgllb;
    DGLLB_REQUIRE(name, "set_option: null name");
    std::call_once(g_opt_once, opt_init);
    for (int k = 0; k < OPT_COUNT; ++k) {
        if (strcmp(kOpts[k].name, name) == 0) {
            int v = 0;
            DGLLB_REQUIRE(opt_parse(kOpts[k], value, &v), "set_option: bad value '%s' for %s", value, name);
            g_opt[k].store(v, std::memory_order_relaxed);
            return DGLLB_OK;
        }
    }
    set_error("set_option: unknown option '%s'", name);
    return DGLLB_ERR_INVALID;
}

int dgllb_get_option(const char* name, int* value_out) {
    using namespace dgllb;
    DGLLB_REQUIRE(name && value_out, "get_option: null pointer");
    for (int k = 0; k < OPT_COUNT; ++k) {
        if (strcmp(kOpts[k].name, name) == 0) {
            *value_out = opt_get(static_cast<Opt>(k));
            return DGLLB_OK;
        }
    }
    set_error("get_option: unknown option '%s'", name);
    return DGLLB_ERR_INVALID;
}

const char* dgllb_last_error(void) { return dgllb::t_err; }

int dgllb_device_info(int* sm_count, int* cc_major, int* cc_minor, int64_t* l2_bytes) {
    dgllb::DevInfo d;
    int rc = dgllb::get_devinfo(&d);
    if (rc != DGLLB_OK) return rc;
    if (sm_count) *sm_count = d.sm_count;
    if (cc_major) *cc_major = d.cc_major;
    if (cc_minor) *cc_minor = d.cc_minor;
    if (l2_bytes) *l2_bytes = d.l2_bytes;
    return DGLLB_OK;
}

int64_t dgllb_launch_count(void) { return dgllb::g_launch_count.load(); }

}  // extern "C"
