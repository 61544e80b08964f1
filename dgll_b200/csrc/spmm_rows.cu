// spmm_rows.cu — whole-row neighbourhood aggregation for short rows (sampled blocks) on sm_100a, optionally
// reading the source table IN PLACE from the shards of a node-range-partitioned table mapped over NVLink.
//
// Replaces, for sum/mean over rows of bounded degree: torch.spmm (dgll/nn/Convolution/gcnconv.py:31), the mean of
// NeighborAggregator (sageconv.py:32-38), DGL's update_all(copy_u, mean) behind SAGEConv on a sampled block
// (GPU Accelerator/CommGNNModel.py:72-77) together with the block's feature gather that feeds it
// (dgll/data/dgraph.py:105, FeatureCache/storage.py:185-209) — the gather is the aggregation's own load.
//
// Why a second kernel next to spmm_rowslab_kernel (spmm.cu): ncu on the headline block (25-edge rows, F=602;
// profiles/r01_spmm_headline.txt) showed the (row, 512-byte slab)-per-warp kernel issue-bound — ~50 warp
// instructions per 512-byte load (per-slab index shuffles, 64-bit c*ldx multiplies, runtime null checks in the
// unrolled loop), 67 % issue-active, 4 dependent rounds of 8 loads per row.  Here:
//   * one warp owns a whole destination row: NS slabs (NS*512 bytes of every source row) per edge, so the column
//     index shuffle and the address arithmetic are paid once per edge, not once per slab;
//   * the lane that loaded col[e] computes the source row's BYTE ADDRESS once (shard lookup or base + c*stride);
//     the address is broadcast with two shuffles, the slab offsets are LDG immediates;
//   * a ROLLING window of D edges (D*NS 128-bit loads per lane) stays in flight across the whole row — no
//     per-round drain;
//   * everything optional (edge values, shards) is a template parameter: the inner loop has no runtime branch
//     besides the row-length test.
// Sharded mode is the B200-native halo exchange fused into the aggregation: shard = id / rows_per_shard
// (node-range partition, SURVEY.md §8 e), remote rows are pulled by the kernel's own LDG.128 over NVSwitch.
// Algorithmic bytes per launch: nnz*(4 + [4] + F*b) + n_dst*(F*4 + r).
#include "common.cuh"
#include "internal.cuh"

namespace dgllb {

struct RowsParams {
    const void* row_ptr;
    int rp64;
    const int* col;
    const float* vals;
    const char* X;                     // single table (shard_ptrs == nullptr)
    long long stride_bytes;            // source row stride in bytes
    const char* const* shard_ptrs;     // device array of shard base pointers (sharded mode)
    int rows_per_shard;
    float* out;
    long long ldo;
    long long n_dst;
    int F;
    int mean;
    const float* row_scale;
    const float* addend;
    long long ld_add;
    const float* bias;
    int epi;
    int out_vec;
    int n_pass;                        // passes over the feature axis (NS slabs each)
};

__device__ __forceinline__ float rows_epi(float v, int epi) {
    if (epi & DGLLB_EPI_RELU) v = fmaxf(v, 0.f);
    if (epi & DGLLB_EPI_ELU) v = v > 0.f ? v : expm1f(v);
    return v;
}

template <typename XT> struct RowsVec;
template <> struct RowsVec<float> {
    static constexpr int A = 4;        // elements per 128-bit load
    typedef float4 raw_t;
    __device__ static __forceinline__ raw_t load(const char* p) { return ldg_nc_f4(reinterpret_cast<const float*>(p)); }
    __device__ static __forceinline__ void unpack(const raw_t& r, float* v) { v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w; }
};
template <> struct RowsVec<__nv_bfloat16> {
    static constexpr int A = 8;
    typedef uint4 raw_t;
    __device__ static __forceinline__ raw_t load(const char* p) { return ldg_nc_u4(p); }
    __device__ static __forceinline__ void unpack(const raw_t& r, float* v) {
        v[0] = bf16lo_to_f32(r.x); v[1] = bf16hi_to_f32(r.x); v[2] = bf16lo_to_f32(r.y); v[3] = bf16hi_to_f32(r.y);
        v[4] = bf16lo_to_f32(r.z); v[5] = bf16hi_to_f32(r.z); v[6] = bf16lo_to_f32(r.w); v[7] = bf16hi_to_f32(r.w);
    }
};

template <typename XT, int NS, int D, bool HAS_VALS, bool SHARDED>
__global__ void __launch_bounds__(128)
spmm_rows_kernel(const RowsParams p) {
    typedef RowsVec<XT> V;
    constexpr int A = V::A;
    constexpr int W = 32 * A;                                   // elements per slab (512 bytes)
    const int lane = threadIdx.x & 31;
    const long long wid = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long row = wid / p.n_pass;
    if (row >= p.n_dst) return;
    const int pass = static_cast<int>(wid - row * p.n_pass);
    const int col0 = pass * NS * W + lane * A;                  // first column of this lane's slab 0
    unsigned on = 0;                                            // bit s: slab s of this lane holds real columns
#pragma unroll
    for (int s = 0; s < NS; ++s) on |= (col0 + s * W < p.F) ? (1u << s) : 0u;
    const long long lane_off = static_cast<long long>(col0) * static_cast<long long>(sizeof(XT));

    long long beg, end;
    if (p.rp64) {
        beg = reinterpret_cast<const long long*>(p.row_ptr)[row];
        end = reinterpret_cast<const long long*>(p.row_ptr)[row + 1];
    } else {
        beg = reinterpret_cast<const int*>(p.row_ptr)[row];
        end = reinterpret_cast<const int*>(p.row_ptr)[row + 1];
    }

    float acc[NS][A];
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int a = 0; a < A; ++a) acc[s][a] = 0.f;

    for (long long e0 = beg; e0 < end; e0 += 32) {
        const int n = static_cast<int>(min(32ll, end - e0));
        // the lane that owns edge e0+lane resolves the source row's byte address once
        const char* my_ptr = nullptr;
        float my_w = 1.f;
        if (lane < n) {
            const int c = __ldg(p.col + e0 + lane);
            if (SHARDED) {
                const int sh = c / p.rows_per_shard;
                my_ptr = p.shard_ptrs[sh] + static_cast<long long>(c - sh * p.rows_per_shard) * p.stride_bytes;
            } else {
                my_ptr = p.X + static_cast<long long>(c) * p.stride_bytes;
            }
            if (HAS_VALS) my_w = __ldg(p.vals + e0 + lane);
        }
        const unsigned long long my_addr = reinterpret_cast<unsigned long long>(my_ptr);

        typename V::raw_t raw[D][NS];
        float w[D];
        // prologue: the first D edges
#pragma unroll
        for (int u = 0; u < D; ++u) {
            if (u < n) {
                const char* src = reinterpret_cast<const char*>(__shfl_sync(0xffffffffu, my_addr, u)) + lane_off;
                if (HAS_VALS) w[u] = __shfl_sync(0xffffffffu, my_w, u);
#pragma unroll
                for (int s = 0; s < NS; ++s)
                    if (on & (1u << s)) raw[u][s] = V::load(src + s * (W * static_cast<int>(sizeof(XT))));
            }
        }
        // steady state: consume edge j, refill its slot with edge j + D
        for (int j0 = 0; j0 < n; j0 += D) {
#pragma unroll
            for (int u = 0; u < D; ++u) {
                const int j = j0 + u;
                if (j < n) {
#pragma unroll
                    for (int s = 0; s < NS; ++s) {
                        if (on & (1u << s)) {
                            float v[A];
                            V::unpack(raw[u][s], v);
#pragma unroll
                            for (int a = 0; a < A; ++a)
                                acc[s][a] = HAS_VALS ? fmaf(w[u], v[a], acc[s][a]) : acc[s][a] + v[a];
                        }
                    }
                    const int jn = j + D;
                    if (jn < n) {
                        const char* src = reinterpret_cast<const char*>(__shfl_sync(0xffffffffu, my_addr, jn)) + lane_off;
                        if (HAS_VALS) w[u] = __shfl_sync(0xffffffffu, my_w, jn);
#pragma unroll
                        for (int s = 0; s < NS; ++s)
                            if (on & (1u << s)) raw[u][s] = V::load(src + s * (W * static_cast<int>(sizeof(XT))));
                    }
                }
            }
        }
    }

    // epilogue: scale / addend / bias / activation / store
    const long long deg = end - beg;
    float scale = 1.f;
    if (p.mean) scale = deg > 0 ? 1.f / static_cast<float>(deg) : 0.f;
    if (p.row_scale) scale *= __ldg(p.row_scale + row);
    const bool plain = !p.addend && !p.bias && !p.epi;
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        if (!(on & (1u << s))) continue;
        const int c0 = col0 + s * W;
        const int valid = min(A, p.F - c0);
        float* __restrict__ o = p.out + row * p.ldo + c0;
#pragma unroll
        for (int a = 0; a < A; ++a) {
            float v = acc[s][a] * scale;
            if (!plain && a < valid) {
                if (p.addend) v += __ldg(p.addend + row * p.ld_add + c0 + a);
                if (p.bias) v += __ldg(p.bias + c0 + a);
                v = rows_epi(v, p.epi);
            }
            acc[s][a] = v;
        }
        if (valid == A && p.out_vec) {
#pragma unroll
            for (int a = 0; a < A; a += 4)
                stg_cs_f4(o + a, make_float4(acc[s][a], acc[s][a + 1], acc[s][a + 2], acc[s][a + 3]));
        } else {
#pragma unroll
            for (int a = 0; a < A; ++a)
                if (a < valid) o[a] = acc[s][a];
        }
    }
}


// ------------------------------------------------------------------------------------------ streaming variant --
// The kernel above pays, PER ROW, a chain of dependent round trips before its window is full (row_ptr -> col -> first
// source rows) and drains the window at the row's end; ncu on the headline block (profiles/r02_spmm_headline.txt):
// 16 resident warps/SM (128 registers), warps active 22 %, DRAM 51 %, every top stall on the first use of a window slot.
// A dependency-free gather of the same shape (tools/membw: gather_rows_2408B, depth 4, 16 warps/SM, 40 % of the rows
// hot in L2) reaches 9.4 TB/s against this kernel's 6.8 — the difference is that per-row start-up and drain.
// Here a warp owns R CONSECUTIVE rows and streams their concatenated edge list: row ends are fetched once (one
// coalesced load, broadcast by shuffle), column ids 32 at a time one chunk ahead, and the window of D source rows
// keeps rolling ACROSS row boundaries — a finished row is scaled and stored while the next row's loads are already
// in flight.  R is chosen so that the grid is about one wave of resident warps.  Plain epilogue only (sum / mean):
// anything else takes the per-row kernel.  Every row is owned by one warp: no atomics, deterministic, CSR edge order.
template <typename XT, int NS, int D, bool HAS_VALS, bool SHARDED>
__global__ void __launch_bounds__(64)      // no register cap: at <= 128 registers the F=602 window spills and loses (0.099 vs 0.087 ms)
spmm_rows_stream_kernel(const RowsParams p, const int R) {
    static_assert(32 % D == 0, "window depth must divide the 32-edge chunk");
    typedef RowsVec<XT> V;
    constexpr int A = V::A;
    constexpr int W = 32 * A;
    const int lane = threadIdx.x & 31;
    const long long wid = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long chunk = wid / p.n_pass;
    const long long r0 = chunk * R;
    if (r0 >= p.n_dst) return;
    const long long r1 = min(r0 + R, p.n_dst);
    const int pass = static_cast<int>(wid - chunk * p.n_pass);
    const int col0 = pass * NS * W + lane * A;
    unsigned on = 0;
#pragma unroll
    for (int s = 0; s < NS; ++s) on |= (col0 + s * W < p.F) ? (1u << s) : 0u;
    const long long lane_off = static_cast<long long>(col0) * static_cast<long long>(sizeof(XT));

    // row boundaries of the owned rows, relative to the first edge: lane l holds the END of row r0 + l (R <= 32)
    long long ebase;
    int my_re = 0x7fffffff;
    {
        const long long ri = min(r0 + lane, r1);   // lanes past the range read row_ptr[r1] (valid), unused
        long long v0, v1;
        if (p.rp64) {
            v0 = reinterpret_cast<const long long*>(p.row_ptr)[r0];
            v1 = reinterpret_cast<const long long*>(p.row_ptr)[min(ri + 1, r1)];
        } else {
            v0 = reinterpret_cast<const int*>(p.row_ptr)[r0];
            v1 = reinterpret_cast<const int*>(p.row_ptr)[min(ri + 1, r1)];
        }
        ebase = v0;
        if (r0 + lane < r1) my_re = static_cast<int>(v1 - v0);
    }
    const int n_rows = static_cast<int>(r1 - r0);
    const int m = __shfl_sync(0xffffffffu, my_re, n_rows - 1);   // edges of the whole range

    const int* __restrict__ cb = p.col + ebase;
    const float* __restrict__ vb = HAS_VALS ? p.vals + ebase : nullptr;
    auto resolve = [&](int e) -> unsigned long long {   // byte address of edge e's source row (0 past the range)
        if (e >= m) return 0ull;
        const int c = __ldg(cb + e);
        const char* ptr;
        if (SHARDED) {
            const int sh = c / p.rows_per_shard;
            ptr = p.shard_ptrs[sh] + static_cast<long long>(c - sh * p.rows_per_shard) * p.stride_bytes;
        } else {
            ptr = p.X + static_cast<long long>(c) * p.stride_bytes;
        }
        return reinterpret_cast<unsigned long long>(ptr);
    };
    unsigned long long my_a = resolve(lane), my_an = resolve(32 + lane);
    float my_w = 1.f, my_wn = 1.f;
    if (HAS_VALS) {
        my_w = lane < m ? __ldg(vb + lane) : 0.f;
        my_wn = 32 + lane < m ? __ldg(vb + 32 + lane) : 0.f;
    }

    int row = 0;                                                  // index inside the owned range
    int rstart = 0, rend = __shfl_sync(0xffffffffu, my_re, 0);

    typename V::raw_t raw[D][NS];
    float w[D];
#pragma unroll
    for (int u = 0; u < D; ++u) {
        const char* src = reinterpret_cast<const char*>(__shfl_sync(0xffffffffu, my_a, u)) + lane_off;
        if (HAS_VALS) w[u] = __shfl_sync(0xffffffffu, my_w, u);
        if (u < m) {
#pragma unroll
            for (int s = 0; s < NS; ++s)
                if (on & (1u << s)) raw[u][s] = V::load(src + s * (W * static_cast<int>(sizeof(XT))));
        }
    }
    float acc[NS][A];
#pragma unroll
    for (int s = 0; s < NS; ++s)
#pragma unroll
        for (int a = 0; a < A; ++a) acc[s][a] = 0.f;

    auto flush = [&]() {   // row (r0 + row) is complete: scale, store, move on
        const int deg = rend - rstart;
        const float scale = p.mean ? (deg > 0 ? 1.f / static_cast<float>(deg) : 0.f) : 1.f;
        float* __restrict__ orow = p.out + (r0 + row) * p.ldo + col0;
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            if (on & (1u << s)) {
                const int valid = min(A, p.F - (col0 + s * W));
                float* __restrict__ o = orow + s * W;
                if (valid == A && p.out_vec) {
#pragma unroll
                    for (int a = 0; a < A; a += 4)
                        stg_cs_f4(o + a, make_float4(acc[s][a] * scale, acc[s][a + 1] * scale, acc[s][a + 2] * scale,
                                                     acc[s][a + 3] * scale));
                } else {
#pragma unroll
                    for (int a = 0; a < A; ++a)
                        if (a < valid) o[a] = acc[s][a] * scale;
                }
            }
#pragma unroll
            for (int a = 0; a < A; ++a) acc[s][a] = 0.f;
        }
        rstart = rend;
        ++row;
        const int nr = __shfl_sync(0xffffffffu, my_re, row & 31);
        rend = row < n_rows ? nr : 0x7fffffff;                    // past the range: nothing ends any more
    };

    for (int jb = 0; jb < m; jb += 32) {
#pragma unroll 1
        for (int k = 0; k < 32; k += D) {
            if (jb + k >= m) break;
            const bool nxt = k + D >= 32;                         // the refills of this group come from the next chunk
            const unsigned long long asel = nxt ? my_an : my_a;
            const float wsel = nxt ? my_wn : my_w;
#pragma unroll
            for (int u = 0; u < D; ++u) {
                const int j = jb + k + u;
                while (j >= rend) flush();                        // rows that ended before edge j (also empty rows)
                if (j < m) {
#pragma unroll
                    for (int s = 0; s < NS; ++s) {
                        if (on & (1u << s)) {
                            float v[A];
                            V::unpack(raw[u][s], v);
#pragma unroll
                            for (int a = 0; a < A; ++a)
                                acc[s][a] = HAS_VALS ? fmaf(w[u], v[a], acc[s][a]) : acc[s][a] + v[a];
                        }
                    }
                }
                const int srcl = (k + D + u) & 31;
                const char* src = reinterpret_cast<const char*>(__shfl_sync(0xffffffffu, asel, srcl)) + lane_off;
                if (HAS_VALS) w[u] = __shfl_sync(0xffffffffu, wsel, srcl);
                if (j + D < m) {
#pragma unroll
                    for (int s = 0; s < NS; ++s)
                        if (on & (1u << s)) raw[u][s] = V::load(src + s * (W * static_cast<int>(sizeof(XT))));
                }
            }
        }
        my_a = my_an;
        my_an = resolve(jb + 64 + lane);
        if (HAS_VALS) {
            my_w = my_wn;
            my_wn = jb + 64 + lane < m ? __ldg(vb + jb + 64 + lane) : 0.f;
        }
    }
    while (row < n_rows) flush();
}

// block size (threads), slabs per warp and window depth can be pinned for A/B runs (options rows_tb/rows_ns/rows_d)
struct RowsTuning { int tb, ns, depth; };
static RowsTuning rows_tuning() { return {opt_get(OPT_ROWS_TB), opt_get(OPT_ROWS_NS), opt_get(OPT_ROWS_D)}; }

// rows per warp of the streaming kernel: 2..16, aiming at about two waves of resident warps (measured on the headline
// block, profiles/r02_spmm_rows_sweep.md: 2-3 rows per warp 87-88 us, per-row kernel 90.8 us, 4-5 rows per warp — about
// ONE wave — 102-122 us: every warp then walks through its start-up chain at the same moment)
static int stream_rows_per_warp(long long n_dst, int n_pass, int sm_count, int resident) {
    const int forced = opt_get(OPT_ROWS_STREAM);
    if (forced > 1) return forced > 32 ? 32 : forced;
    const double slots = static_cast<double>(sm_count) * resident * 2.0;
    long long r = static_cast<long long>(static_cast<double>(n_dst) * n_pass / slots);
    if (r < 2) r = 2;
    if (r > 16) r = 16;
    return static_cast<int>(r);
}

// The sharded aggregation runs on a graph branch BESIDE the training step; its warps sit on NVLink round trips (~2,000
// cycles) while holding registers, and at full residency they take ~94 % of the register file, so the training branch's
// CTAs cannot start on an SM until it drains.  Option rows_sharded_bps = n caps the resident 64-thread blocks per SM at n
// by asking for 1/n of the shared memory the kernel never touches (NVLink needs ~2 MB in flight per GPU: 10 warps per SM
// with an 8-deep window hold 6 MB).
static size_t sharded_throttle_smem(bool sharded) {
    if (!sharded) return 0;
    const int bps = opt_get(OPT_ROWS_SHARDED_BPS);
    if (bps < 5 || bps > 16) return 0;
    size_t b = (static_cast<size_t>(227) * 1024 / bps - 1024) & ~static_cast<size_t>(1023);
    return b > 48 * 1024 ? 48 * 1024 : b;
}

template <typename XT, int NS, int D, bool HAS_VALS, bool SHARDED>
static int rows_launch(const RowsParams& p, int tb, cudaStream_t st) {
    const size_t thr_smem = sharded_throttle_smem(SHARDED);
    constexpr int DS = (D >= 8) ? 8 : (D >= 3 ? 4 : 2);                       // streaming depth must divide 32
    const bool plain = !p.row_scale && !p.addend && !p.bias && !p.epi;
    // bf16 rows: the streaming variant measured slower than the per-row kernel (70.5 vs 78.3 us) — fp32 only
    if (plain && opt_get(OPT_ROWS_STREAM) != 1 && (sizeof(XT) == 4 || opt_get(OPT_ROWS_STREAM) > 1) && p.n_dst >= 256) {
        DevInfo di;
        int rc = get_devinfo(&di);
        if (rc != DGLLB_OK) return rc;
        static const int resident = [] {   // resident warps per SM of this instantiation (64-thread blocks)
            int nb = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, spmm_rows_stream_kernel<XT, NS, DS, HAS_VALS, SHARDED>, 64, 0) != cudaSuccess || nb <= 0)
                nb = 8;
            return nb * 2;
        }();
        int res_eff = resident;
        if (thr_smem) { const int cap = opt_get(OPT_ROWS_SHARDED_BPS) * 2; if (cap < res_eff) res_eff = cap; }
        const int R = stream_rows_per_warp(p.n_dst, p.n_pass, di.sm_count, res_eff);
        if (R >= 2) {
            const long long warps = (p.n_dst + R - 1) / R * p.n_pass;
            const long long blocks = (warps + 1) / 2;
            DGLLB_REQUIRE(blocks < (1ll << 31), "spmm_rows: grid too large (%lld blocks)", blocks);
            spmm_rows_stream_kernel<XT, NS, DS, HAS_VALS, SHARDED><<<static_cast<unsigned>(blocks), 64, thr_smem, st>>>(p, R);
            DGLLB_LAUNCH_CHECK();
            return DGLLB_OK;
        }
    }
    const long long warps = p.n_dst * p.n_pass;
    const long long blocks = (warps * 32 + tb - 1) / tb;
    DGLLB_REQUIRE(blocks < (1ll << 31), "spmm_rows: grid too large (%lld blocks)", blocks);
    spmm_rows_kernel<XT, NS, D, HAS_VALS, SHARDED><<<static_cast<unsigned>(blocks), tb, thr_smem, st>>>(p);
    DGLLB_LAUNCH_CHECK();
    return DGLLB_OK;
}

template <typename XT, int NS, int D, bool SHARDED>
static int rows_launch_v(const RowsParams& p, int tb, cudaStream_t st) {
    return p.vals ? rows_launch<XT, NS, D, true, SHARDED>(p, tb, st) : rows_launch<XT, NS, D, false, SHARDED>(p, tb, st);
}

// NS slabs per warp and the window depth that goes with it (registers: NS*D*4 for the window + NS*A accumulators).
// Measured on the headline block (F=602 fp32, 25-edge rows) and on F=128/256 blocks: profiles/r02_spmm_rows_sweep.md.
template <typename XT, bool SHARDED>
static int rows_dispatch(RowsParams& p, cudaStream_t st) {
    constexpr int A = RowsVec<XT>::A;
    const int n_slabs = (p.F + 32 * A - 1) / (32 * A);
    const RowsTuning t = rows_tuning();
    int max_ns = (t.ns >= 1 && t.ns <= 5) ? t.ns : (A == 4 ? 5 : 3);
    p.n_pass = (n_slabs + max_ns - 1) / max_ns;
    const int ns = (n_slabs + p.n_pass - 1) / p.n_pass;
    const int tb = (t.tb == 32 || t.tb == 64 || t.tb == 128) ? t.tb : 64;
    // rows behind NVLink come back ~3x later than local HBM rows (peer LDG ~2,000 cycles): a deeper window when the rows
    // are narrow enough for the registers
    const int d = t.depth ? t.depth : ((SHARDED && ns <= 2) ? 8 : 0);
    switch (ns) {
        case 1: return d == 2 ? rows_launch_v<XT, 1, 2, SHARDED>(p, tb, st)
                     : d == 8 ? rows_launch_v<XT, 1, 8, SHARDED>(p, tb, st)
                              : rows_launch_v<XT, 1, 4, SHARDED>(p, tb, st);
        case 2: return d == 2 ? rows_launch_v<XT, 2, 2, SHARDED>(p, tb, st)
                     : d == 8 ? rows_launch_v<XT, 2, 8, SHARDED>(p, tb, st)
                              : rows_launch_v<XT, 2, 4, SHARDED>(p, tb, st);
        case 3: return d == 2 ? rows_launch_v<XT, 3, 2, SHARDED>(p, tb, st)
                     : d == 5 ? rows_launch_v<XT, 3, 5, SHARDED>(p, tb, st)
                              : rows_launch_v<XT, 3, 3, SHARDED>(p, tb, st);
        case 4: return d == 2 ? rows_launch_v<XT, 4, 2, SHARDED>(p, tb, st)
                     : d == 4 ? rows_launch_v<XT, 4, 4, SHARDED>(p, tb, st)
                              : rows_launch_v<XT, 4, 3, SHARDED>(p, tb, st);
        default: return d == 2 ? rows_launch_v<XT, 5, 2, SHARDED>(p, tb, st)
                      : d == 3 ? rows_launch_v<XT, 5, 3, SHARDED>(p, tb, st)
                               : rows_launch_v<XT, 5, 4, SHARDED>(p, tb, st);
    }
}

// Entry used by spmm_run (spmm.cu) for short-row sum/mean aggregation over one table.
int spmm_rows_try(const SpmmParams& sp, int x_dtype, cudaStream_t st) {
    const int esz = x_dtype == DGLLB_F32 ? 4 : 2;
    const int A = 16 / esz;
    if (!sp.col || sp.row_cnt || sp.argmax || sp.F <= 16 * A) return DGLLB_ERR_UNSUPPORTED;
    if (!aligned16(sp.X) || (sp.ldx % A) != 0) return DGLLB_ERR_UNSUPPORTED;
    RowsParams p;
    p.row_ptr = sp.row_ptr; p.rp64 = sp.rp64; p.col = sp.col; p.vals = sp.vals;
    p.X = static_cast<const char*>(sp.X); p.stride_bytes = sp.ldx * esz;
    p.shard_ptrs = nullptr; p.rows_per_shard = 1;
    p.out = sp.out; p.ldo = sp.ldo; p.n_dst = sp.n_dst; p.F = sp.F; p.mean = sp.mean;
    p.row_scale = sp.row_scale; p.addend = sp.addend; p.ld_add = sp.ld_add; p.bias = sp.bias; p.epi = sp.epi;
    p.out_vec = aligned16(sp.out) && (sp.ldo % 4 == 0);
    p.n_pass = 1;
    return x_dtype == DGLLB_F32 ? rows_dispatch<float, false>(p, st) : rows_dispatch<__nv_bfloat16, false>(p, st);
}

}  // namespace dgllb

using namespace dgllb;

extern "C" int dgllb_spmm_csr_sharded(const void* row_ptr, int row_ptr_is64, const int32_t* col_idx,
                                      const float* values, const void* const* shard_ptrs, int n_shards,
                                      int64_t rows_per_shard, int64_t stride_bytes, int x_dtype, float* out,
                                      int64_t ldo, int64_t n_dst, int F, int reduce, void* stream) {
    DGLLB_REQUIRE(n_dst >= 0 && F >= 0, "spmm_sharded: negative size");
    if (n_dst == 0 || F == 0) return DGLLB_OK;
    DGLLB_REQUIRE(row_ptr && col_idx && shard_ptrs && out, "spmm_sharded: null pointer");
    DGLLB_REQUIRE(n_shards >= 1 && rows_per_shard >= 1 && rows_per_shard < (1ll << 31),
                  "spmm_sharded: bad shard geometry");
    DGLLB_REQUIRE(reduce == DGLLB_SUM || reduce == DGLLB_MEAN, "spmm_sharded: reduce must be sum or mean");
    DGLLB_REQUIRE(x_dtype == DGLLB_F32 || x_dtype == DGLLB_BF16, "spmm_sharded: unknown dtype %d", x_dtype);
    const int esz = x_dtype == DGLLB_F32 ? 4 : 2;
    DGLLB_REQUIRE(stride_bytes % 16 == 0 && stride_bytes >= static_cast<int64_t>(F) * esz,
                  "spmm_sharded: shard rows must be 16-byte multiples holding F elements (stride %lld, F %d)",
                  (long long)stride_bytes, F);
    DGLLB_REQUIRE(ldo >= F && n_dst < (1ll << 31), "spmm_sharded: bad output shape");
    RowsParams p;
    p.row_ptr = row_ptr; p.rp64 = row_ptr_is64; p.col = col_idx; p.vals = values;
    p.X = nullptr; p.stride_bytes = stride_bytes;
    p.shard_ptrs = reinterpret_cast<const char* const*>(shard_ptrs);
    p.rows_per_shard = static_cast<int>(rows_per_shard);
    p.out = out; p.ldo = ldo; p.n_dst = n_dst; p.F = F; p.mean = reduce == DGLLB_MEAN;
    p.row_scale = nullptr; p.addend = nullptr; p.ld_add = 0; p.bias = nullptr; p.epi = 0;
    p.out_vec = aligned16(out) && (ldo % 4 == 0);
    p.n_pass = 1;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return x_dtype == DGLLB_F32 ? rows_dispatch<float, true>(p, st) : rows_dispatch<__nv_bfloat16, true>(p, st);
}
