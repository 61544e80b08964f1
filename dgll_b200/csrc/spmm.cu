// spmm.cu — CSR neighbourhood aggregation (SpMM) for sm_100a.
//
// Replaces: torch.spmm (dgll/nn/Convolution/gcnconv.py:31), torch.sparse.mm
// (Evaluation/PPI/gcn_model.py:76), the edge loop of gcn_fused_kernel.cu:41-57,
// NeighborAggregator mean/sum/max (sageconv.py:32-38), scatter() pooling
// (GlobalPooling/Pooling.py:37,59,81) and the DGL update_all SpMM behind
// GraphConv/SAGEConv (GPU Accelerator/CommGNNModel.py:23-28,72-77).
//
// Design (HBM-bound gather, no tensor cores):
//   work item  = (destination row, feature slab); a slab is LANES x 16 bytes of
//                one feature row, so one LDG.128 per lane fetches a contiguous
//                512-byte (LANES=32) piece of a source row — fully coalesced;
//   row-split  : a group of LANES lanes owns an item and walks the row's edges;
//                column indices are loaded LANES at a time (coalesced) and
//                broadcast with shuffles; U=8 source rows are in flight per lane;
//   nnz-split  : with a plan, rows longer than `chunk` edges are cut into
//                chunks that run first (longest work first) and combine through
//                128-bit RED.ADD into pre-zeroed rows, then a finalize kernel
//                applies the epilogue;
//   reductions : fp32 accumulation in CSR edge order (sum/mean), max with argmax.
// Algorithmic bytes per launch: nnz*(4 + [4 values] + F*b) + n_dst*(F*4 + r).
#include "common.cuh"
#include "internal.cuh"

namespace dgllb {

__device__ __forceinline__ long long load_rp(const void* p, int is64, long long i) {
    return is64 ? reinterpret_cast<const long long*>(p)[i]
                : static_cast<long long>(reinterpret_cast<const int*>(p)[i]);
}

__device__ __forceinline__ float apply_epi(float v, int epi) {
    if (epi & DGLLB_EPI_RELU) v = fmaxf(v, 0.f);
    if (epi & DGLLB_EPI_ELU) v = v > 0.f ? v : expm1f(v);
    return v;
}

// ---------------------------------------------------------------- traits --
// XT = float : VEC elements per lane = 4 (16 B) or 1
// XT = __nv_bfloat16 : VEC elements per lane = 8 (16 B) or 1
template <typename XT, int VE>
struct RowLoad;

template <>
struct RowLoad<float, 4> {
    static constexpr int kAcc = 4;
    typedef float4 raw_t;
    __device__ static __forceinline__ raw_t load(const float* p) { return ldg_nc_f4(p); }
    __device__ static __forceinline__ raw_t zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
    __device__ static __forceinline__ void unpack(const raw_t& r, float* v) {
        v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
    }
};
template <>
struct RowLoad<float, 1> {
    static constexpr int kAcc = 1;
    typedef float raw_t;
    __device__ static __forceinline__ raw_t load(const float* p) { return ldg_nc_f1(p); }
    __device__ static __forceinline__ raw_t zero() { return 0.f; }
    __device__ static __forceinline__ void unpack(const raw_t& r, float* v) { v[0] = r; }
};
template <>
struct RowLoad<__nv_bfloat16, 8> {
    static constexpr int kAcc = 8;
    typedef uint4 raw_t;
    __device__ static __forceinline__ raw_t load(const __nv_bfloat16* p) { return ldg_nc_u4(p); }
    __device__ static __forceinline__ raw_t zero() { return make_uint4(0u, 0u, 0u, 0u); }
    __device__ static __forceinline__ void unpack(const raw_t& r, float* v) {
        v[0] = bf16lo_to_f32(r.x); v[1] = bf16hi_to_f32(r.x);
        v[2] = bf16lo_to_f32(r.y); v[3] = bf16hi_to_f32(r.y);
        v[4] = bf16lo_to_f32(r.z); v[5] = bf16hi_to_f32(r.z);
        v[6] = bf16lo_to_f32(r.w); v[7] = bf16hi_to_f32(r.w);
    }
};
template <>
struct RowLoad<__nv_bfloat16, 1> {
    static constexpr int kAcc = 1;
    typedef unsigned short raw_t;
    __device__ static __forceinline__ raw_t load(const __nv_bfloat16* p) {
        return __ldg(reinterpret_cast<const unsigned short*>(p));
    }
    __device__ static __forceinline__ raw_t zero() { return 0; }
    __device__ static __forceinline__ void unpack(const raw_t& r, float* v) {
        v[0] = __uint_as_float(static_cast<uint32_t>(r) << 16);
    }
};

constexpr int kThreads = 256;
constexpr int kUnroll = 8;

// scale / addend / bias / activation / store of one finished row piece (A consecutive columns at col0)
template <int A>
__device__ __forceinline__ void store_row(const SpmmParams& p, long long row, int col0, float* acc, long long deg) {
    float scale = 1.f;
    if (p.mean) scale = deg > 0 ? 1.f / static_cast<float>(deg) : 0.f;
    if (p.row_scale) scale *= __ldg(p.row_scale + row);
    float* __restrict__ o = p.out + row * p.ldo + col0;
    const int valid = min(A, p.F - col0);
#pragma unroll
    for (int a = 0; a < A; ++a) {
        float v = acc[a] * scale;
        if (a < valid) {
            if (p.addend) v += __ldg(p.addend + row * p.ld_add + col0 + a);
            if (p.bias) v += __ldg(p.bias + col0 + a);
            v = apply_epi(v, p.epi);
        }
        acc[a] = v;
    }
    if (A >= 4 && valid == A && p.out_vec) {
#pragma unroll
        for (int a = 0; a < A; a += 4)
            stg_cs_f4(o + a, make_float4(acc[a], acc[a + 1], acc[a + 2], acc[a + 3]));
    } else {
#pragma unroll
        for (int a = 0; a < A; ++a)
            if (a < valid) o[a] = acc[a];
    }
}


// ------------------------------------------------------------ main kernel --
template <typename XT, int VE, int LANES, bool IS_MAX>
__global__ void __launch_bounds__(kThreads)
spmm_rowslab_kernel(const SpmmParams p) {
    typedef RowLoad<XT, VE> L;
    constexpr int A = L::kAcc;
    const int lig = threadIdx.x & (LANES - 1);                 // lane in group
    const long long group = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) / LANES;
    const unsigned gmask = (LANES == 32) ? 0xffffffffu
                                         : (((1u << LANES) - 1u) << ((threadIdx.x & 31) & ~(LANES - 1)));

    const long long n_heavy_groups = p.n_heavy_items * p.n_slabs;
    long long row, beg, end, deg;
    int slab;
    bool heavy = false;
    long long it = 0;
    if (group < n_heavy_groups) {
        it = group / p.n_slabs;
        slab = static_cast<int>(group - it * p.n_slabs);
        const int2 hc = p.heavy_items[it];
        row = hc.x;
        const long long rb = load_rp(p.row_ptr, p.rp64, row);
        const long long re = load_rp(p.row_ptr, p.rp64, row + 1);
        deg = re - rb;
        beg = rb + static_cast<long long>(hc.y) * p.chunk_edges;
        end = min(re, beg + p.chunk_edges);
        heavy = true;
    } else {
        const long long g2 = group - n_heavy_groups;
        row = g2 / p.n_slabs;
        if (row >= p.n_dst) return;
        slab = static_cast<int>(g2 - row * p.n_slabs);
        beg = load_rp(p.row_ptr, p.rp64, row);
        end = load_rp(p.row_ptr, p.rp64, row + 1);
        if (p.row_cnt) end = min(end, beg + max(0, __ldg(p.row_cnt + row)));  // legacy num_neighbors guard
        deg = end - beg;
        if (!IS_MAX && p.chunk_edges > 0 && deg > p.chunk_edges) return;  // done by heavy items
    }

    const int col0 = (slab * LANES + lig) * A;  // first feature column of this lane
    const bool lane_on = col0 < p.F;
    const XT* __restrict__ Xb = reinterpret_cast<const XT*>(p.X) + col0;

    float acc[A];
    int amax[IS_MAX ? A : 1];
#pragma unroll
    for (int a = 0; a < A; ++a) acc[a] = IS_MAX ? -INFINITY : 0.f;
    if (IS_MAX) {
#pragma unroll
        for (int a = 0; a < A; ++a) amax[a] = -1;
    }

    for (long long e0 = beg; e0 < end; e0 += LANES) {
        const int n = static_cast<int>(min(static_cast<long long>(LANES), end - e0));
        long long my_c = 0;
        float my_w = 1.f;
        if (lig < n) {
            my_c = p.col ? static_cast<long long>(__ldg(p.col + e0 + lig)) : (e0 + lig);
            if (p.vals) my_w = __ldg(p.vals + e0 + lig);
        }
        for (int k = 0; k < n; k += kUnroll) {
            typename L::raw_t raw[kUnroll];
            float w[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int src = (k + u) & (LANES - 1);
                // identity-column mode can exceed int32; shuffle the 64-bit index as two halves
                long long c;
                if (p.col) {
                    c = static_cast<long long>(__shfl_sync(gmask, static_cast<int>(my_c), src, LANES));
                } else {
                    c = __shfl_sync(gmask, my_c, src, LANES);
                }
                w[u] = __shfl_sync(gmask, my_w, src, LANES);
                if (k + u < n && lane_on) {
                    raw[u] = L::load(Xb + c * p.ldx);
                } else {
                    raw[u] = L::zero();
                    w[u] = IS_MAX ? NAN : 0.f;  // NAN marks "no edge" for max
                }
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                float v[A];
                L::unpack(raw[u], v);
                if (IS_MAX) {
                    if (w[u] == w[u]) {  // not the NAN marker
#pragma unroll
                        for (int a = 0; a < A; ++a) {
                            const float t = w[u] * v[a];
                            if (t > acc[a]) {
                                acc[a] = t;
                                amax[a] = static_cast<int>(e0 + k + u);
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int a = 0; a < A; ++a) acc[a] = fmaf(w[u], v[a], acc[a]);
                }
            }
        }
    }

    if (!lane_on) return;

    float scale = 1.f;
    if (p.mean) scale = deg > 0 ? 1.f / static_cast<float>(deg) : 0.f;
    if (p.row_scale) scale *= __ldg(p.row_scale + row);
    if (IS_MAX && deg == 0) {
#pragma unroll
        for (int a = 0; a < A; ++a) acc[a] = 0.f;
    }
#pragma unroll
    for (int a = 0; a < A; ++a) acc[a] *= scale;

    float* __restrict__ o = p.out + row * p.ldo + col0;
    const int valid = min(A, p.F - col0);

    if (heavy) {
        // partial sums of a split row go to the item's workspace row; spmm_finalize_heavy_kernel adds the chunks of a
        // row in chunk order and runs the epilogue (round 1 used RED.ADD here: heavy-row sums were not reproducible)
        float* __restrict__ w = p.heavy_ws + it * p.ld_hws + col0;
        if (A % 4 == 0 && valid == A) {
#pragma unroll
            for (int a = 0; a < A; a += 4)
                *reinterpret_cast<float4*>(w + a) = make_float4(acc[a], acc[a + 1], acc[a + 2], acc[a + 3]);
        } else {
#pragma unroll
            for (int a = 0; a < A; ++a)
                if (a < valid) w[a] = acc[a];
        }
        return;
    }

#pragma unroll
    for (int a = 0; a < A; ++a) {
        if (a < valid) {
            float v = acc[a];
            if (p.addend) v += __ldg(p.addend + row * p.ld_add + col0 + a);
            if (p.bias) v += __ldg(p.bias + col0 + a);
            acc[a] = apply_epi(v, p.epi);
        }
    }
    if (A >= 4 && valid == A && p.out_vec) {
#pragma unroll
        for (int a = 0; a < A; a += 4)
            stg_cs_f4(o + a, make_float4(acc[a], acc[a + 1], acc[a + 2], acc[a + 3]));
    } else {
#pragma unroll
        for (int a = 0; a < A; ++a)
            if (a < valid) o[a] = acc[a];
    }
    if (IS_MAX && p.argmax) {
        int* am = p.argmax + row * static_cast<long long>(p.F) + col0;
#pragma unroll
        for (int a = 0; a < A; ++a)
            if (a < valid) am[a] = amax[a];
    }
}

// ------------------------------------------------------- streaming kernel --
// Row-aligned nnz-split ("merge-path" by edges, boundaries snapped to row starts): warp (q, slab) owns the
// rows whose first edge lies in [q*T, (q+1)*T) and streams their concatenated edge list with a ROLLING
// window of kDepth source rows in flight per lane: as soon as edge j's piece is consumed, the load for edge
// j+kDepth is issued, across row boundaries, so short rows (sampled blocks: <= fanout edges) never drain the
// memory pipeline the way one-warp-per-row does.  Every row is owned by exactly one warp: no atomics,
// deterministic, fp32 accumulation in CSR edge order.  Column indices (and edge values) are fetched 32 at a
// time, one chunk ahead, and broadcast with shuffles; row ends travel the same way.
constexpr int kDepth = 8;

template <typename XT, int VE, bool HAS_VALS>
__global__ void __launch_bounds__(kThreads, 3)
spmm_stream_kernel(const SpmmParams p, const int T, const long long n_chunks) {
    typedef RowLoad<XT, VE> L;
    constexpr int A = L::kAcc;
    const int lane = threadIdx.x & 31;
    const long long wid = (static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x) >> 5;
    const long long q = wid / p.n_slabs;
    if (q >= n_chunks) return;
    const int slab = static_cast<int>(wid - q * p.n_slabs);

    // first row whose start is >= t: even lanes search q*T, odd lanes (q+1)*T
    long long r0, r1;
    {
        const long long t = (q + (lane & 1)) * static_cast<long long>(T);
        long long lo = 0, hi = p.n_dst;
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (load_rp(p.row_ptr, p.rp64, mid) < t) lo = mid + 1; else hi = mid;
        }
        r0 = __shfl_sync(0xffffffffu, lo, 0);
        r1 = __shfl_sync(0xffffffffu, lo, 1);
        if (q == n_chunks - 1) r1 = p.n_dst;
    }
    if (r0 >= r1) return;  // this edge range lies inside one long row owned by an earlier warp
    const long long ebase = load_rp(p.row_ptr, p.rp64, r0);
    const int m = static_cast<int>(load_rp(p.row_ptr, p.rp64, r1) - ebase);

    const int col0 = (slab * 32 + lane) * A;
    const bool lane_on = col0 < p.F;
    const XT* __restrict__ Xb = reinterpret_cast<const XT*>(p.X) + col0;
    const long long ldx = p.ldx;
    const int* __restrict__ cb = p.col + ebase;
    const float* __restrict__ vb = HAS_VALS ? p.vals + ebase : nullptr;

    int my_c = lane < m ? __ldg(cb + lane) : 0;
    int my_cn = 32 + lane < m ? __ldg(cb + 32 + lane) : 0;
    float my_w = 1.f, my_wn = 1.f;
    if (HAS_VALS) {
        my_w = lane < m ? __ldg(vb + lane) : 0.f;
        my_wn = 32 + lane < m ? __ldg(vb + 32 + lane) : 0.f;
    }
    // row-end offsets (relative to ebase) of rows row_w0 + lane; INT_MAX past the owned range
    long long row = r0, row_w0 = r0;
    int my_re = (row_w0 + lane < r1) ? static_cast<int>(load_rp(p.row_ptr, p.rp64, row_w0 + lane + 1) - ebase)
                                     : 0x7fffffff;
    int rstart = 0;
    int rend = __shfl_sync(0xffffffffu, my_re, 0);

    typename L::raw_t raw[kDepth];
    float w[kDepth];
#pragma unroll
    for (int u = 0; u < kDepth; ++u) {
        const int c = __shfl_sync(0xffffffffu, my_c, u);
        if (HAS_VALS) w[u] = __shfl_sync(0xffffffffu, my_w, u);
        if (lane_on && u < m) raw[u] = L::load(Xb + static_cast<long long>(c) * ldx);
    }
    float acc[A];
#pragma unroll
    for (int a = 0; a < A; ++a) acc[a] = 0.f;

    // The flush inside the rolling loop stays tiny (no calls, no spills): it stores scale * sum only.  The rare
    // epilogue terms (row_scale / addend / bias / activation) are applied by a second pass of the SAME thread
    // over the rows it has just written (they are still in L1/L2; n_dst*F*4 B is small next to the gather).
    const bool need_epi = p.row_scale || p.addend || p.bias || p.epi;
    const int valid = min(A, p.F - col0);
    const bool full_vec = p.out_vec && valid == A;
    auto flush = [&]() {
        if (lane_on) {
            const int deg = rend - rstart;
            const float scale = p.mean ? (deg > 0 ? 1.f / static_cast<float>(deg) : 0.f) : 1.f;
            float* __restrict__ o = p.out + row * p.ldo + col0;
            if (full_vec) {
#pragma unroll
                for (int a = 0; a < A; a += 4) {
                    const float4 v4 = make_float4(acc[a] * scale, acc[a + 1] * scale, acc[a + 2] * scale,
                                                  acc[a + 3] * scale);
                    if (need_epi) *reinterpret_cast<float4*>(o + a) = v4; else stg_cs_f4(o + a, v4);
                }
            } else {
#pragma unroll
                for (int a = 0; a < A; ++a)
                    if (a < valid) o[a] = acc[a] * scale;
            }
        }
#pragma unroll
        for (int a = 0; a < A; ++a) acc[a] = 0.f;
        rstart = rend;
        ++row;
        if (row - row_w0 == 32) {
            row_w0 = row;
            my_re = (row_w0 + lane < r1)
                        ? static_cast<int>(load_rp(p.row_ptr, p.rp64, row_w0 + lane + 1) - ebase)
                        : 0x7fffffff;
        }
        rend = __shfl_sync(0xffffffffu, my_re, static_cast<int>(row - row_w0));
    };

    for (int jb = 0; jb < m; jb += 32) {
#pragma unroll 1
        for (int k8 = 0; k8 < 32; k8 += kDepth) {
            if (jb + k8 >= m) break;
            // the refills of this group read indices kn = k8+kDepth.. : current chunk or the next one (warp-uniform)
            const bool nxt = k8 + kDepth >= 32;
            const int csel = nxt ? my_cn : my_c;
            const float wsel = nxt ? my_wn : my_w;
#pragma unroll
            for (int u = 0; u < kDepth; ++u) {
                const int j = jb + k8 + u;
                while (j >= rend) flush();  // rows that ended before edge j (also empty rows)
                if (j < m) {
                    float v[A];
                    L::unpack(raw[u], v);
#pragma unroll
                    for (int a = 0; a < A; ++a) acc[a] = HAS_VALS ? fmaf(w[u], v[a], acc[a]) : acc[a] + v[a];
                }
                // refill the slot with edge j + kDepth
                const int src = (k8 + kDepth + u) & 31;
                const int c = __shfl_sync(0xffffffffu, csel, src);
                if (HAS_VALS) w[u] = __shfl_sync(0xffffffffu, wsel, src);
                if (lane_on && j + kDepth < m) raw[u] = L::load(Xb + static_cast<long long>(c) * ldx);
            }
        }
        my_c = my_cn;
        my_cn = jb + 64 + lane < m ? __ldg(cb + jb + 64 + lane) : 0;
        if (HAS_VALS) {
            my_w = my_wn;
            my_wn = jb + 64 + lane < m ? __ldg(vb + jb + 64 + lane) : 0.f;
        }
    }
    while (row < r1) flush();

    if (need_epi && lane_on) {
        for (long long r = r0; r < r1; ++r) {
            float* __restrict__ o = p.out + r * p.ldo + col0;
            const float rs = p.row_scale ? __ldg(p.row_scale + r) : 1.f;
#pragma unroll
            for (int a = 0; a < A; ++a) {
                if (a < valid) {
                    float v = o[a] * rs;
                    if (p.addend) v += __ldg(p.addend + r * p.ld_add + col0 + a);
                    if (p.bias) v += __ldg(p.bias + col0 + a);
                    o[a] = apply_epi(v, p.epi);
                }
            }
        }
    }
}

// merge the partial rows of every split row: one block per k == 0 item (its chunks are items [w, w + nc)), fixed order
__global__ void spmm_finalize_heavy_kernel(const SpmmParams p) {
    const long long w = blockIdx.x;
    if (w >= p.n_heavy_items) return;
    const int2 it = p.heavy_items[w];
    if (it.y != 0) return;
    const long long row = it.x;
    const long long deg = load_rp(p.row_ptr, p.rp64, row + 1) - load_rp(p.row_ptr, p.rp64, row);
    const int nc = static_cast<int>((deg + p.chunk_edges - 1) / p.chunk_edges);
    float* o = p.out + row * p.ldo;
    const float* base = p.heavy_ws + w * p.ld_hws;
    for (int f = threadIdx.x; f < p.F; f += blockDim.x) {
        float v = 0.f;
        for (int c = 0; c < nc; ++c) v += base[static_cast<long long>(c) * p.ld_hws + f];
        if (p.addend) v += p.addend[row * p.ld_add + f];
        if (p.bias) v += p.bias[f];
        o[f] = apply_epi(v, p.epi);
    }
}

// ------------------------------------------------------------------ plan --
__global__ void plan_count_kernel(const void* row_ptr, int rp64, long long n_rows, int chunk,
                                  unsigned long long* counters) {
    const long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const long long deg = load_rp(row_ptr, rp64, r + 1) - load_rp(row_ptr, rp64, r);
    if (deg > chunk) {
        atomicAdd(&counters[0], 1ull);
        atomicAdd(&counters[1], static_cast<unsigned long long>((deg + chunk - 1) / chunk));
    }
}

__global__ void plan_fill_kernel(const void* row_ptr, int rp64, long long n_rows, int chunk,
                                 unsigned long long* counters, int* heavy_rows, int2* items) {
    const long long r = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r >= n_rows) return;
    const long long deg = load_rp(row_ptr, rp64, r + 1) - load_rp(row_ptr, rp64, r);
    if (deg > chunk) {
        const unsigned long long hi = atomicAdd(&counters[2], 1ull);
        heavy_rows[hi] = static_cast<int>(r);
        const int nc = static_cast<int>((deg + chunk - 1) / chunk);
        const unsigned long long base = atomicAdd(&counters[3], static_cast<unsigned long long>(nc));
        for (int k = 0; k < nc; ++k) items[base + k] = make_int2(static_cast<int>(r), k);
    }
}

}  // namespace dgllb


using namespace dgllb;

extern "C" int dgllb_csr_plan_create(const void* row_ptr, int row_ptr_is64, int64_t n_rows,
                                     int chunk_edges, void* stream, dgllb_csr_plan** plan_out) {
    DGLLB_REQUIRE(row_ptr && plan_out, "csr_plan_create: null pointer");
    DGLLB_REQUIRE(n_rows >= 0 && n_rows < (1ll << 31), "csr_plan_create: n_rows out of range");
    if (chunk_edges <= 0) chunk_edges = 4096;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned long long* d_cnt = nullptr;
    DGLLB_CUDA_TRY(cudaMalloc(&d_cnt, 4 * sizeof(unsigned long long)));
    cudaError_t e = cudaMemsetAsync(d_cnt, 0, 4 * sizeof(unsigned long long), st);
    unsigned long long h_cnt[4] = {0, 0, 0, 0};
    const int tb = 256;
    const unsigned grid = static_cast<unsigned>((n_rows + tb - 1) / tb);
    if (e == cudaSuccess && n_rows > 0) {
        plan_count_kernel<<<grid, tb, 0, st>>>(row_ptr, row_ptr_is64, n_rows, chunk_edges, d_cnt);
        g_launch_count.fetch_add(1);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        cudaFree(d_cnt);
        set_error("csr_plan_create: %s", cudaGetErrorString(e));
        return DGLLB_ERR_CUDA;
    }
    dgllb_csr_plan* pl = new dgllb_csr_plan();
    pl->chunk_edges = chunk_edges;
    pl->n_rows = n_rows;
    pl->n_heavy_rows = static_cast<long long>(h_cnt[0]);
    pl->n_items = static_cast<long long>(h_cnt[1]);
    pl->heavy_rows = nullptr;
    pl->items = nullptr;
    if (pl->n_heavy_rows > 0) {
        e = cudaMalloc(&pl->heavy_rows, pl->n_heavy_rows * sizeof(int));
        if (e == cudaSuccess) e = cudaMalloc(&pl->items, pl->n_items * sizeof(int2));
        if (e == cudaSuccess) {
            plan_fill_kernel<<<grid, tb, 0, st>>>(row_ptr, row_ptr_is64, n_rows, chunk_edges, d_cnt,
                                                  pl->heavy_rows, pl->items);
            g_launch_count.fetch_add(1);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            cudaFree(d_cnt);
            dgllb_csr_plan_destroy(pl);
            set_error("csr_plan_create: %s", cudaGetErrorString(e));
            return DGLLB_ERR_CUDA;
        }
    }
    cudaFree(d_cnt);
    *plan_out = pl;
    return DGLLB_OK;
}

extern "C" int dgllb_csr_plan_info(const dgllb_csr_plan* plan, int64_t* n_heavy_rows,
                                   int64_t* n_chunks, int* chunk_edges) {
    DGLLB_REQUIRE(plan, "csr_plan_info: null plan");
    if (n_heavy_rows) *n_heavy_rows = plan->n_heavy_rows;
    if (n_chunks) *n_chunks = plan->n_items;
    if (chunk_edges) *chunk_edges = plan->chunk_edges;
    return DGLLB_OK;
}

extern "C" void dgllb_csr_plan_destroy(dgllb_csr_plan* plan) {
    if (!plan) return;
    if (plan->heavy_rows) cudaFree(plan->heavy_rows);
    if (plan->items) cudaFree(plan->items);
    delete plan;
}

// --------------------------------------------------------------- dispatch --
template <typename XT, int VE, int LANES>
static int launch_spmm(const SpmmParams& p, bool is_max, cudaStream_t st) {
    const long long groups = (p.n_heavy_items + p.n_dst) * p.n_slabs;
    if (groups == 0) return DGLLB_OK;
    // 64-thread blocks: a block's registers are held until its slowest row finishes, so small blocks retire evenly on
    // ragged rows (headline block 0.127 -> 0.115 ms per step, full graphs 8 % faster at F=256; profiles/r01_block_size.md)
    const int tb_env = opt_get(OPT_SPMM_TB);
    const int tb = (tb_env == 32 || tb_env == 64 || tb_env == 128 || tb_env == 256) ? tb_env : 64;
    const int gpb = tb / LANES;
    const long long blocks = (groups + gpb - 1) / gpb;
    DGLLB_REQUIRE(blocks < (1ll << 31), "spmm: grid too large (%lld blocks)", blocks);
    if (is_max)
        spmm_rowslab_kernel<XT, VE, LANES, true><<<static_cast<unsigned>(blocks), tb, 0, st>>>(p);
    else
        spmm_rowslab_kernel<XT, VE, LANES, false><<<static_cast<unsigned>(blocks), tb, 0, st>>>(p);
    DGLLB_LAUNCH_CHECK();
    return DGLLB_OK;
}

// edges per warp for the streaming kernel: enough warps for ~4 full waves, at least 32, at most 1024 edges
static int stream_chunk_edges(long long nnz, int n_slabs, int sm_count) {
    const long long target_warps = static_cast<long long>(sm_count) * 32 * 4;
    long long t = (nnz * n_slabs + target_warps - 1) / target_warps;
    t = (t + 31) / 32 * 32;
    if (t < 32) t = 32;
    if (t > 1024) t = 1024;
    return static_cast<int>(t);
}

template <typename XT, int VE>
static int launch_spmm_stream(SpmmParams& p, long long nnz, cudaStream_t st) {
    DevInfo di;
    int rc = get_devinfo(&di);
    if (rc != DGLLB_OK) return rc;
    p.n_slabs = (p.F + 32 * VE - 1) / (32 * VE);
    const int T = stream_chunk_edges(nnz, p.n_slabs, di.sm_count);
    const long long n_chunks = nnz > 0 ? (nnz + T - 1) / T : 1;
    const long long warps = n_chunks * p.n_slabs;
    const long long blocks = (warps * 32 + kThreads - 1) / kThreads;
    DGLLB_REQUIRE(blocks < (1ll << 31), "spmm: grid too large (%lld blocks)", blocks);
    if (p.vals)
        spmm_stream_kernel<XT, VE, true><<<static_cast<unsigned>(blocks), kThreads, 0, st>>>(p, T, n_chunks);
    else
        spmm_stream_kernel<XT, VE, false><<<static_cast<unsigned>(blocks), kThreads, 0, st>>>(p, T, n_chunks);
    DGLLB_LAUNCH_CHECK();
    return DGLLB_OK;
}

template <typename XT, int VE>
static int launch_spmm_lanes(SpmmParams& p, bool is_max, cudaStream_t st) {
    // smallest group whose slab covers F, so narrow rows do not idle lanes
    const int per_lane = VE;
    if (VE > 1 && p.F <= 8 * per_lane) {
        p.n_slabs = 1;
        return launch_spmm<XT, VE, 8>(p, is_max, st);
    }
    if (VE > 1 && p.F <= 16 * per_lane) {
        p.n_slabs = 1;
        return launch_spmm<XT, VE, 16>(p, is_max, st);
    }
    p.n_slabs = (p.F + 32 * per_lane - 1) / (32 * per_lane);
    return launch_spmm<XT, VE, 32>(p, is_max, st);
}

namespace dgllb {
int spmm_run(SpmmParams& p, int x_dtype, bool is_max, const dgllb_csr_plan* plan, cudaStream_t st) {
    const bool use_plan = plan && !is_max && plan->n_heavy_rows > 0 && !p.row_cnt;
    if (use_plan) {
        p.heavy_items = plan->items;
        p.n_heavy_items = plan->n_items;
        p.chunk_edges = plan->chunk_edges;
        p.ld_hws = (static_cast<long long>(p.F) + 7) / 8 * 8;
        DGLLB_REQUIRE(plan->n_items < (1ll << 31), "spmm: too many plan items");
        DGLLB_CUDA_TRY(cudaMallocAsync(&p.heavy_ws, sizeof(float) * static_cast<size_t>(plan->n_items) * p.ld_hws, st));
    }
    int rc;
    p.out_vec = aligned16(p.out) && (p.ldo % 4 == 0);
    // streaming kernel: sum/mean over an explicit col_idx with a host-known nnz bound, wide rows, no split plan
    bool stream_ok = !is_max && !use_plan && p.col && !p.row_cnt && p.nnz_hint >= 0;
    // Kernel choice (measured on B200, profiles/r01_kernels*.jsonl, profiles/r02_spmm_rows_sweep.md): the whole-row
    // rolling-window kernel (spmm_rows.cu) on sampled blocks, the row-aligned streaming kernel on large inputs
    // (>= 2^19 edges), row-split + nnz-split plan on skewed full graphs.  Option spmm_kernel (env DGLLB_SPMM_KERNEL,
    // dgllb_set_option) = rowsplit|stream|wholerow pins one family for A/B runs.
    const int force = opt_get(OPT_SPMM_KERNEL);
    const bool rows_ok = !is_max && !use_plan && (force == 0 ? (p.nnz_hint >= 0 && p.nnz_hint < (1ll << 19)) : force == 3);
    if (rows_ok) {
        rc = spmm_rows_try(p, x_dtype, st);
        if (rc != DGLLB_ERR_UNSUPPORTED) return rc;
    }
    if (force != 0 && force != 2) stream_ok = false;
    if (force == 0 && p.nnz_hint < (1ll << 19)) stream_ok = false;
    if (x_dtype == DGLLB_F32) {
        // 128-bit LOADS need only the source table aligned; the output falls back to scalar stores by itself
        const bool vec = aligned16(p.X) && (p.ldx % 4 == 0);
        if (vec && stream_ok && p.F > 16 * 4) rc = launch_spmm_stream<float, 4>(p, p.nnz_hint, st);
        else rc = vec ? launch_spmm_lanes<float, 4>(p, is_max, st) : launch_spmm_lanes<float, 1>(p, is_max, st);
    } else {
        const bool vec = aligned16(p.X) && (p.ldx % 8 == 0);
        if (vec && stream_ok && p.F > 16 * 8) rc = launch_spmm_stream<__nv_bfloat16, 8>(p, p.nnz_hint, st);
        else rc = vec ? launch_spmm_lanes<__nv_bfloat16, 8>(p, is_max, st)
                      : launch_spmm_lanes<__nv_bfloat16, 1>(p, is_max, st);
    }
    if (use_plan) {
        if (rc == DGLLB_OK) {
            spmm_finalize_heavy_kernel<<<static_cast<unsigned>(plan->n_items), 128, 0, st>>>(p);
            g_launch_count.fetch_add(1, std::memory_order_relaxed);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) { set_error("spmm: %s", cudaGetErrorString(e)); rc = DGLLB_ERR_CUDA; }
        }
        cudaFreeAsync(p.heavy_ws, st);
    }
    return rc;
}
}  // namespace dgllb

extern "C" int dgllb_spmm_csr(const void* row_ptr, int row_ptr_is64, const int32_t* col_idx,
                              const float* values, const void* X, int x_dtype, int64_t ldx,
                              float* out, int64_t ldo, int64_t n_dst, int64_t n_src, int64_t nnz,
                              int F, int reduce, const float* row_scale, const float* addend,
                              int64_t ld_add, const float* bias, int epilogue,
                              int32_t* argmax_out, const dgllb_csr_plan* plan, void* stream) {
    DGLLB_REQUIRE(n_dst >= 0 && n_src >= 0 && F >= 0, "spmm: negative size");
    if (n_dst == 0 || F == 0) return DGLLB_OK;
    DGLLB_REQUIRE(row_ptr && X && out, "spmm: null pointer");
    DGLLB_REQUIRE(ldx >= F && ldo >= F, "spmm: leading dimension smaller than F (ldx=%lld ldo=%lld F=%d)",
                  (long long)ldx, (long long)ldo, F);
    DGLLB_REQUIRE(reduce == DGLLB_SUM || reduce == DGLLB_MEAN || reduce == DGLLB_MAX,
                  "spmm: unknown reduce %d", reduce);
    DGLLB_REQUIRE(x_dtype == DGLLB_F32 || x_dtype == DGLLB_BF16, "spmm: unknown dtype %d", x_dtype);
    DGLLB_REQUIRE(!addend || ld_add >= F, "spmm: ld_add smaller than F");
    DGLLB_REQUIRE(n_dst < (1ll << 31), "spmm: n_dst must fit int32");
    DGLLB_REQUIRE(!plan || plan->n_rows == n_dst, "spmm: plan was built for %lld rows, got %lld",
                  plan ? plan->n_rows : 0ll, (long long)n_dst);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SpmmParams p;
    p.row_ptr = row_ptr;
    p.rp64 = row_ptr_is64;
    p.col = col_idx;
    p.vals = values;
    p.X = X;
    p.ldx = ldx;
    p.out = out;
    p.ldo = ldo;
    p.n_dst = n_dst;
    p.F = F;
    p.n_slabs = 1;
    p.mean = reduce == DGLLB_MEAN;
    p.row_scale = row_scale;
    p.addend = addend;
    p.ld_add = ld_add;
    p.bias = bias;
    p.epi = epilogue;
    p.argmax = argmax_out;
    p.heavy_items = nullptr;
    p.n_heavy_items = 0;
    p.chunk_edges = 0;
    p.heavy_ws = nullptr;
    p.ld_hws = 0;
    p.row_cnt = nullptr;
    p.nnz_hint = nnz;
    return spmm_run(p, x_dtype, reduce == DGLLB_MAX, plan, st);
}

// ----------------------------------------------------------------- SDDMM --
namespace dgllb {

// one warp per row; the row of A stays in registers (F <= 32*4*kMaxChunks) and
// each edge's row of B streams through with 128-bit loads.
template <int VE>
__global__ void __launch_bounds__(256)
sddmm_kernel(const void* row_ptr, int rp64, const int* __restrict__ col, const float* __restrict__ A,
             long long lda, const float* __restrict__ B, long long ldb, float* __restrict__ out_e,
             long long n_rows, int F) {
    const int lane = threadIdx.x & 31;
    const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (row >= n_rows) return;
    const long long beg = load_rp(row_ptr, rp64, row), end = load_rp(row_ptr, rp64, row + 1);
    const float* a = A + row * lda;
    for (long long e = beg; e < end; ++e) {
        const float* b = B + static_cast<long long>(__ldg(col + e)) * ldb;
        float s = 0.f;
        if (VE == 4) {
            for (int f = lane * 4; f < F; f += 128) {
                if (f + 4 <= F) {
                    const float4 x = __ldg(reinterpret_cast<const float4*>(a + f));
                    const float4 y = ldg_nc_f4(b + f);
                    s = fmaf(x.x, y.x, s); s = fmaf(x.y, y.y, s);
                    s = fmaf(x.z, y.z, s); s = fmaf(x.w, y.w, s);
                } else {
                    for (int t = f; t < F; ++t) s = fmaf(__ldg(a + t), __ldg(b + t), s);
                }
            }
        } else {
            for (int f = lane; f < F; f += 32) s = fmaf(__ldg(a + f), __ldg(b + f), s);
        }
        s = warp_sum(s);
        if (lane == 0) out_e[e] = s;
    }
}

__global__ void max_backward_kernel(const int* __restrict__ col, const int* __restrict__ argmax,
                                    const float* __restrict__ g, long long ldg, float* gx,
                                    long long ldx, long long n_dst, int F) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n_dst * F) return;
    const long long r = i / F;
    const int f = static_cast<int>(i - r * F);
    const int e = argmax[i];
    if (e >= 0) atomicAdd(gx + static_cast<long long>(col ? col[e] : e) * ldx + f, g[r * ldg + f]);
}

}  // namespace dgllb

extern "C" int dgllb_sddmm_csr(const void* row_ptr, int row_ptr_is64, const int32_t* col_idx,
                               const float* A, int64_t lda, const float* B, int64_t ldb,
                               float* out_e, int64_t n_rows, int F, void* stream) {
    if (n_rows == 0) return DGLLB_OK;
    DGLLB_REQUIRE(row_ptr && col_idx && A && B && out_e, "sddmm: null pointer");
    DGLLB_REQUIRE(lda >= F && ldb >= F && F >= 0, "sddmm: bad leading dimension");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long blocks = (n_rows * 32 + 255) / 256;
    DGLLB_REQUIRE(blocks < (1ll << 31), "sddmm: grid too large");
    const bool vec = aligned16(A) && aligned16(B) && lda % 4 == 0 && ldb % 4 == 0;
    if (vec)
        sddmm_kernel<4><<<static_cast<unsigned>(blocks), 256, 0, st>>>(row_ptr, row_ptr_is64, col_idx, A, lda,
                                                                       B, ldb, out_e, n_rows, F);
    else
        sddmm_kernel<1><<<static_cast<unsigned>(blocks), 256, 0, st>>>(row_ptr, row_ptr_is64, col_idx, A, lda,
                                                                       B, ldb, out_e, n_rows, F);
    DGLLB_LAUNCH_CHECK();
    return DGLLB_OK;
}

extern "C" int dgllb_spmm_max_backward(const int32_t* col_idx, const int32_t* argmax,
                                       const float* grad_out, int64_t ldg, float* grad_X,
                                       int64_t ldx, int64_t n_dst, int F, void* stream) {
    if (n_dst == 0 || F == 0) return DGLLB_OK;
    DGLLB_REQUIRE(argmax && grad_out && grad_X, "max_backward: null pointer");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long total = n_dst * F;
    const long long blocks = (total + 255) / 256;
    DGLLB_REQUIRE(blocks < (1ll << 31), "max_backward: grid too large");
    max_backward_kernel<<<static_cast<unsigned>(blocks), 256, 0, st>>>(col_idx, argmax, grad_out, ldg, grad_X,
                                                                       ldx, n_dst, F);
    DGLLB_LAUNCH_CHECK();
    return DGLLB_OK;
}
