// gat.cu — fused SDDMM + edge-softmax + aggregation for multi-head GAT (sm_100a).
//
// Replaces sparseGatConv.forward (dgll/nn/Convolution/gatconv.py:111-148: edge
// gather+cat, exp(-leakyrelu), two COO SpMMs, divide) and the dense N^2 masked
// softmax of gatConv.forward (:30-54), plus their autograd (SpecialSpmmFunction
// backward :71-81, which materialises a dense N x N product).
//
// GAT's attention logit is additive: z_ij = lrelu(a_l.Wh_i + a_r.Wh_j) = lrelu(el_i + er_j),
// so the SDDMM collapses to two per-node scalars per head (computed by the dense
// transform's caller) and the whole layer is ONE pass over the edges:
//   work item = (dst row, head, slab of the head's D columns); a group of LANES
//   lanes walks the row's edges LANES at a time: each lane scores one edge
//   (gathers er_j), the group does an online-softmax update (running max m,
//   running sum l, accumulator rescale), then the chunk's source rows stream
//   through with 128-bit loads, 8 in flight per lane.  No [E] tensor is written.
// Backward (deterministic, no atomics): pass 1 over the CSR recomputes alpha,
// forms dalpha = <g_i, Wh_j> (the SDDMM) and writes dz per edge; pass 2 over the
// transposed CSR accumulates d_Wh_j = sum_i alpha_ij g_i and d_er_j.
// Algorithmic bytes fwd: nnz*(4 + F*4 + heads*4) + n_dst*(F*4 + heads*4 + r).
#include "common.cuh"
#include "internal.cuh"
#include <stdlib.h>

namespace dgllb {

struct GatParams {
    const void* row_ptr;
    int rp64;
    const int* col;
    const float* Wh;
    long long ldw;
    const float* el;
    const float* er;
    long long ld_e;
    float* out;
    long long ldo;
    float* row_max;
    float* row_sum;
    long long n_dst;
    int heads;
    int D;
    int n_slabs;
    float slope;
    float sign;  // +1 softmax(lrelu), -1 exp(-lrelu)/sum
    int epi;
    // attention dropout (gatconv.py:37 / :132): keep iff hash(seed, edge, head) >= drop_thresh; kept weights * drop_scale
    unsigned drop_thresh;  // 0 = no dropout
    float drop_scale;
    unsigned long long seed;
    // nnz-split of long rows (whole-row kernel only): partial (acc, m, l) per (row, chunk) item in `ws`
    int chunk;            // 0 = no split
    const int2* items;
    long long n_items;
    float* ws;            // [n_items, heads*D + 8]: acc, then m[h] at +FD+h and l[h] at +FD+heads+h
};

__device__ __forceinline__ long long gat_rp(const void* p, int is64, long long i) {
    return is64 ? reinterpret_cast<const long long*>(p)[i]
                : static_cast<long long>(reinterpret_cast<const int*>(p)[i]);
}

template <int LANES>
__device__ __forceinline__ float group_max(float v, unsigned gmask) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(gmask, v, o, LANES));
    return v;
}
template <int LANES>
__device__ __forceinline__ float group_sum(float v, unsigned gmask) {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o, LANES);
    return v;
}

// counter-based dropout mask: one decision per (edge position in the CSR, head), reproducible in the backward
__device__ __forceinline__ float gat_keep_scale(unsigned long long seed, long long e, int h, int H, unsigned thresh,
                                                float scale) {
    unsigned long long z = seed + static_cast<unsigned long long>(e) * static_cast<unsigned>(H) + static_cast<unsigned>(h);
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return static_cast<unsigned>(z >> 32) >= thresh ? scale : 0.f;
}

constexpr int kGatThreads = 256;
constexpr int kGatUnroll = 8;

template <int VE, int LANES>
__global__ void __launch_bounds__(kGatThreads)
gat_forward_kernel(const GatParams p) {
    const int lig = threadIdx.x & (LANES - 1);
    const long long group = (static_cast<long long>(blockIdx.x) * kGatThreads + threadIdx.x) / LANES;
    const unsigned gmask = (LANES == 32) ? 0xffffffffu
                                         : (((1u << LANES) - 1u) << ((threadIdx.x & 31) & ~(LANES - 1)));
    const int per_row = p.heads * p.n_slabs;
    const long long row = group / per_row;
    if (row >= p.n_dst) return;
    const int rem = static_cast<int>(group - row * per_row);
    const int head = rem / p.n_slabs;
    const int slab = rem - head * p.n_slabs;
    const long long beg = gat_rp(p.row_ptr, p.rp64, row), end = gat_rp(p.row_ptr, p.rp64, row + 1);

    const int dcol = (slab * LANES + lig) * VE;  // column inside the head
    const bool lane_on = dcol < p.D;
    const int col0 = head * p.D + dcol;
    const float* __restrict__ Wb = p.Wh + col0;
    const float el_i = __ldg(p.el + row * p.ld_e + head);

    float m = -INFINITY, l = 0.f;
    float acc[VE];
#pragma unroll
    for (int a = 0; a < VE; ++a) acc[a] = 0.f;

    // software pipeline over the row's LANES-edge chunks: column ids are fetched two chunks ahead and the source
    // attention scalars er[col] one chunk ahead, so the dependent chain col -> er -> exp -> row loads of the next
    // chunk overlaps the feature-row loads of the current one (profiles/r01_gat_forward.txt: latency bound).
    const float* __restrict__ erh = p.er + head;
    int c_cur = 0, c_nxt = 0;
    float er_cur = 0.f;
    if (beg + lig < end) {
        c_cur = __ldg(p.col + beg + lig);
        er_cur = __ldg(erh + static_cast<long long>(c_cur) * p.ld_e);
    }
    if (beg + LANES + lig < end) c_nxt = __ldg(p.col + beg + LANES + lig);

    for (long long e0 = beg; e0 < end; e0 += LANES) {
        const int n = static_cast<int>(min(static_cast<long long>(LANES), end - e0));
        // prefetch: er of the next chunk (its ids arrived during the previous iteration), ids of the one after
        float er_nxt = 0.f;
        if (e0 + LANES + lig < end) er_nxt = __ldg(erh + static_cast<long long>(c_nxt) * p.ld_e);
        int c_nxt2 = 0;
        if (e0 + 2 * LANES + lig < end) c_nxt2 = __ldg(p.col + e0 + 2 * LANES + lig);
        const int my_c = c_cur;
        float my_s = -INFINITY;
        if (lig < n) {
            float z = el_i + er_cur;
            z = z > 0.f ? z : p.slope * z;
            my_s = p.sign * z;
        }
        c_cur = c_nxt;
        er_cur = er_nxt;
        c_nxt = c_nxt2;
        const float mc = group_max<LANES>(my_s, gmask);
        const float m_new = fmaxf(m, mc);
        const float corr = (m == -INFINITY) ? 0.f : expf(m - m_new);
        float my_p = (lig < n) ? expf(my_s - m_new) : 0.f;
        l = l * corr + group_sum<LANES>(my_p, gmask);
        // dropout acts on the normalised attention: the row sum keeps every edge, the aggregation drops some
        if (p.drop_thresh) my_p *= gat_keep_scale(p.seed, e0 + lig, head, p.heads, p.drop_thresh, p.drop_scale);
        m = m_new;
#pragma unroll
        for (int a = 0; a < VE; ++a) acc[a] *= corr;

        for (int k = 0; k < n; k += kGatUnroll) {
            float4 raw[kGatUnroll];
            float w[kGatUnroll];
#pragma unroll
            for (int u = 0; u < kGatUnroll; ++u) {
                const int src = (k + u) & (LANES - 1);
                const long long c = __shfl_sync(gmask, my_c, src, LANES);
                w[u] = __shfl_sync(gmask, my_p, src, LANES);
                if (k + u < n && lane_on) {
                    if (VE == 4) raw[u] = ldg_nc_f4(Wb + c * p.ldw);
                    else raw[u].x = ldg_nc_f1(Wb + c * p.ldw);
                } else {
                    raw[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    w[u] = 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < kGatUnroll; ++u) {
                acc[0] = fmaf(w[u], raw[u].x, acc[0]);
                if (VE == 4) {
                    acc[1] = fmaf(w[u], raw[u].y, acc[1]);
                    acc[2] = fmaf(w[u], raw[u].z, acc[2]);
                    acc[3] = fmaf(w[u], raw[u].w, acc[3]);
                }
            }
        }
    }

    if (slab == 0 && lig == 0) {
        if (p.row_max) p.row_max[row * p.heads + head] = m;
        if (p.row_sum) p.row_sum[row * p.heads + head] = l;
    }
    if (!lane_on) return;
    const float inv = l > 0.f ? 1.f / l : 0.f;
    float* __restrict__ o = p.out + row * p.ldo + col0;
    const int valid = min(VE, p.D - dcol);
#pragma unroll
    for (int a = 0; a < VE; ++a) {
        float v = acc[a] * inv;
        if (p.epi & DGLLB_EPI_RELU) v = fmaxf(v, 0.f);
        if (p.epi & DGLLB_EPI_ELU) v = v > 0.f ? v : expm1f(v);
        acc[a] = v;
    }
    if (VE == 4 && valid == 4) {
        stg_cs_f4(o, make_float4(acc[0], acc[1], acc[2], acc[3]));
    } else {
#pragma unroll
        for (int a = 0; a < VE; ++a)
            if (a < valid) o[a] = acc[a];
    }
}


// ------------------------------------------------- whole-row forward kernel --
// One WARP per destination row, all heads at once (heads <= 4, heads*D <= 128*NV floats, D % 4 == 0): lane l owns the
// 16-byte vectors l, l+32, ... of the concatenated head outputs, so a source row (e.g. 4 x 64 floats = 1 KB) is
// fetched by NV coalesced LDG.128 per lane and its column id / score work is done ONCE per edge instead of once per
// (edge, head) as in gat_forward_kernel (profiles/r01_gat_forward.txt: 11.0 G warp instructions, latency bound).
// Per 32-edge chunk: lane e scores edge e for every head (er[col] fetched as one vector), the chunk's softmax
// statistics are 2 warp reductions per head, scores and column ids are parked in shared memory and every lane reads
// back the weight of ITS head with a broadcast LDS.  Column ids / er values of the next chunks are prefetched.
// Launched with WARPS warps per block (template parameter): 1 by default — a block's registers are released when its
// slowest row finishes, so one-warp blocks retire evenly on ragged rows (profiles/r01_block_size.md).

// HEAVY = false: warp w owns row w (rows longer than p.chunk edges are skipped when a plan is given).
// HEAVY = true : warp w owns plan item w = (row, k): edges [k*chunk, (k+1)*chunk) of a long row; it writes its partial
//                online-softmax state (unnormalised acc, running max m, running sum l) to p.ws and
//                gat_combine_heavy_kernel merges the chunks of a row (exact log-sum-exp merge, no atomics).
template <int NV, bool HEAVY, int U, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
gat_forward_row_kernel(const GatParams p) {
    __shared__ int s_c[WARPS][32];
    __shared__ float s_p[WARPS][32][4];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long wid = static_cast<long long>(blockIdx.x) * WARPS + warp;
    const int H = p.heads, FD = p.heads * p.D;
    long long row, beg, end;
    if (HEAVY) {
        if (wid >= p.n_items) return;
        const int2 it = p.items[wid];
        row = it.x;
        const long long rb = gat_rp(p.row_ptr, p.rp64, row), re = gat_rp(p.row_ptr, p.rp64, row + 1);
        beg = rb + static_cast<long long>(it.y) * p.chunk;
        end = min(re, beg + p.chunk);
    } else {
        row = wid;
        if (row >= p.n_dst) return;
        beg = gat_rp(p.row_ptr, p.rp64, row);
        end = gat_rp(p.row_ptr, p.rp64, row + 1);
        if (p.chunk > 0 && end - beg > p.chunk) return;  // done by the heavy items
    }

    // this lane's vectors and the head each one belongs to
    int hsel[NV];
    bool von[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c0 = (i * 32 + lane) * 4;
        von[i] = c0 < FD;
        hsel[i] = von[i] ? c0 / p.D : 0;
    }
    float el_i[4], m[4], l[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        el_i[h] = h < H ? __ldg(p.el + row * p.ld_e + h) : 0.f;
        m[h] = -INFINITY;
        l[h] = 0.f;
    }
    float acc[NV][4];
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int a = 0; a < 4; ++a) acc[i][a] = 0.f;

    auto load_er = [&](int c, float* dst) {
#pragma unroll
        for (int h = 0; h < 4; ++h) dst[h] = h < H ? __ldg(p.er + static_cast<long long>(c) * p.ld_e + h) : 0.f;
    };
    int c_cur = 0, c_nxt = 0;
    float er_cur[4] = {0.f, 0.f, 0.f, 0.f};
    if (beg + lane < end) {
        c_cur = __ldg(p.col + beg + lane);
        load_er(c_cur, er_cur);
    }
    if (beg + 32 + lane < end) c_nxt = __ldg(p.col + beg + 32 + lane);

    for (long long e0 = beg; e0 < end; e0 += 32) {
        const int n = static_cast<int>(min(32ll, end - e0));
        float er_nxt[4] = {0.f, 0.f, 0.f, 0.f};
        if (e0 + 32 + lane < end) load_er(c_nxt, er_nxt);
        int c_nxt2 = 0;
        if (e0 + 64 + lane < end) c_nxt2 = __ldg(p.col + e0 + 64 + lane);

        float corr[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            corr[h] = 1.f;
            if (h < H) {
                float sc = -INFINITY;
                if (lane < n) {
                    float z = el_i[h] + er_cur[h];
                    z = z > 0.f ? z : p.slope * z;
                    sc = p.sign * z;
                }
                const float m_new = fmaxf(m[h], warp_max(sc));
                corr[h] = (m[h] == -INFINITY) ? 0.f : expf(m[h] - m_new);
                const float pe = (lane < n) ? expf(sc - m_new) : 0.f;
                l[h] = l[h] * corr[h] + warp_sum(pe);
                m[h] = m_new;
                s_p[warp][lane][h] = p.drop_thresh ? pe * gat_keep_scale(p.seed, e0 + lane, h, H, p.drop_thresh, p.drop_scale)
                                                   : pe;
            }
        }
        s_c[warp][lane] = c_cur;
        __syncwarp();
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const float cr = hsel[i] == 0 ? corr[0] : hsel[i] == 1 ? corr[1] : hsel[i] == 2 ? corr[2] : corr[3];
#pragma unroll
            for (int a = 0; a < 4; ++a) acc[i][a] *= cr;
        }
        for (int k = 0; k < n; k += U) {
            float4 raw[U][NV];
            float w[U][NV];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool on = k + u < n;
                const int c = s_c[warp][(k + u) & 31];
                const float* src = p.Wh + static_cast<long long>(c) * p.ldw + lane * 4;
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    if (on && von[i]) {
                        raw[u][i] = ldg_nc_f4(src + i * 128);
                        w[u][i] = s_p[warp][(k + u) & 31][hsel[i]];
                    } else {
                        raw[u][i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        w[u][i] = 0.f;
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int i = 0; i < NV; ++i) {
                    acc[i][0] = fmaf(w[u][i], raw[u][i].x, acc[i][0]);
                    acc[i][1] = fmaf(w[u][i], raw[u][i].y, acc[i][1]);
                    acc[i][2] = fmaf(w[u][i], raw[u][i].z, acc[i][2]);
                    acc[i][3] = fmaf(w[u][i], raw[u][i].w, acc[i][3]);
                }
        }
        __syncwarp();  // everyone is done with s_c / s_p before the next chunk overwrites them
        c_cur = c_nxt;
#pragma unroll
        for (int h = 0; h < 4; ++h) er_cur[h] = er_nxt[h];
        c_nxt = c_nxt2;
    }

    if (HEAVY) {
        float* w = p.ws + wid * (FD + 8);  // row of FD floats + (m, l) per head, padded to keep 16-byte alignment
        if (lane < H) {
            float mm = m[0], ll = l[0];
#pragma unroll
            for (int h = 1; h < 4; ++h)
                if (lane == h) { mm = m[h]; ll = l[h]; }
            w[FD + lane] = mm;
            w[FD + H + lane] = ll;
        }
#pragma unroll
        for (int i = 0; i < NV; ++i)
            if (von[i])
                *reinterpret_cast<float4*>(w + (i * 32 + lane) * 4) =
                    make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        return;
    }
    if (lane < H) {
        float mm = m[0], ll = l[0];
#pragma unroll
        for (int h = 1; h < 4; ++h)
            if (lane == h) { mm = m[h]; ll = l[h]; }
        if (p.row_max) p.row_max[row * H + lane] = mm;
        if (p.row_sum) p.row_sum[row * H + lane] = ll;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        if (!von[i]) continue;
        const float lh = hsel[i] == 0 ? l[0] : hsel[i] == 1 ? l[1] : hsel[i] == 2 ? l[2] : l[3];
        const float inv = lh > 0.f ? 1.f / lh : 0.f;
        float v[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            float t = acc[i][a] * inv;
            if (p.epi & DGLLB_EPI_RELU) t = fmaxf(t, 0.f);
            if (p.epi & DGLLB_EPI_ELU) t = t > 0.f ? t : expm1f(t);
            v[a] = t;
        }
        stg_cs_f4(p.out + row * p.ldo + (i * 32 + lane) * 4, make_float4(v[0], v[1], v[2], v[3]));
    }
}

// merge the chunk partials of every heavy row: warp = the k == 0 item of a row; its chunks are items [w, w + nc)
__global__ void __launch_bounds__(256)
gat_combine_heavy_kernel(const GatParams p) {
    const int lane = threadIdx.x & 31;
    const long long wid = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (wid >= p.n_items) return;
    const int2 it = p.items[wid];
    if (it.y != 0) return;
    const long long row = it.x;
    const int H = p.heads, FD = p.heads * p.D, W = FD + 8;
    const long long deg = gat_rp(p.row_ptr, p.rp64, row + 1) - gat_rp(p.row_ptr, p.rp64, row);
    const int nc = static_cast<int>((deg + p.chunk - 1) / p.chunk);
    const float* base = p.ws + wid * W;
    for (int c0 = lane * 4; c0 < FD; c0 += 128) {
        const int h = c0 / p.D;
        float M = -INFINITY;
        for (int c = 0; c < nc; ++c) M = fmaxf(M, base[static_cast<long long>(c) * W + FD + h]);
        float L = 0.f;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int c = 0; c < nc; ++c) {
            const float* wc = base + static_cast<long long>(c) * W;
            const float mc = wc[FD + h];
            const float sc = mc == -INFINITY ? 0.f : expf(mc - M);
            L += wc[FD + H + h] * sc;
            const float4 v = *reinterpret_cast<const float4*>(wc + c0);
            a.x = fmaf(v.x, sc, a.x); a.y = fmaf(v.y, sc, a.y); a.z = fmaf(v.z, sc, a.z); a.w = fmaf(v.w, sc, a.w);
        }
        const float inv = L > 0.f ? 1.f / L : 0.f;
        float v[4] = {a.x * inv, a.y * inv, a.z * inv, a.w * inv};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (p.epi & DGLLB_EPI_RELU) v[k] = fmaxf(v[k], 0.f);
            if (p.epi & DGLLB_EPI_ELU) v[k] = v[k] > 0.f ? v[k] : expm1f(v[k]);
        }
        *reinterpret_cast<float4*>(p.out + row * p.ldo + c0) = make_float4(v[0], v[1], v[2], v[3]);
        if ((c0 % p.D) == 0) {  // first vector of head h records the row statistics
            if (p.row_max) p.row_max[row * H + h] = M;
            if (p.row_sum) p.row_sum[row * H + h] = L;
        }
    }
}

// ------------------------------------------------------------- backward --
struct GatBwdParams {
    const void* row_ptr;
    int rp64;
    const int* col;
    const void* t_row_ptr;
    const int* t_col;
    const int* perm;
    const float* Wh;
    long long ldw;
    const float* el;
    const float* er;
    long long ld_e;
    const float* out;
    long long ldo;
    const float* row_max;
    const float* row_sum;
    const float* g;
    long long ldg;
    float* d_Wh;
    long long ldd;
    float* d_el;
    float* d_er;
    long long ld_de;
    float2* ws;  // (alpha, dz) per (edge, head)
    long long n_dst;
    long long n_src;
    int heads;
    int D;
    float slope;
    float sign;
    unsigned drop_thresh;  // same mask as the forward (seed, edge position, head)
    float drop_scale;
    unsigned long long seed;
    // nnz-split of long rows: pass 1 over the forward CSR (e_*), pass 2 over the transposed CSR (t_*).
    // Partial results go to workspaces and are merged by gat_bwd_combine_* (deterministic, no atomics).
    int e_chunk;
    const int2* e_items;
    long long e_n_items;
    float* e_ws;          // [e_n_items, heads]        partial d_el
    int t_chunk;
    const int2* t_items;
    long long t_n_items;
    float* t_ws;          // [t_n_items, heads*D + heads]  partial d_Wh rows, then partial d_er
};

// Pass 1: group = (dst row i, head).  Lane columns: chunk k covers (k*LANES + lig)*VE.
template <int VE, int LANES, int NCH, bool HEAVY>
__global__ void __launch_bounds__(kGatThreads)
gat_backward_edge_kernel(const GatBwdParams p) {
    const int lig = threadIdx.x & (LANES - 1);
    const long long group = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) / LANES;
    const unsigned gmask = (LANES == 32) ? 0xffffffffu
                                         : (((1u << LANES) - 1u) << ((threadIdx.x & 31) & ~(LANES - 1)));
    const long long unit = group / p.heads;       // dst row, or plan item (row, chunk) when HEAVY
    const int head = static_cast<int>(group - unit * p.heads);
    long long row, beg, end;
    if (HEAVY) {
        if (unit >= p.e_n_items) return;
        const int2 it = p.e_items[unit];
        row = it.x;
        const long long rb = gat_rp(p.row_ptr, p.rp64, row), re = gat_rp(p.row_ptr, p.rp64, row + 1);
        beg = rb + static_cast<long long>(it.y) * p.e_chunk;
        end = min(re, beg + p.e_chunk);
    } else {
        row = unit;
        if (row >= p.n_dst) return;
        beg = gat_rp(p.row_ptr, p.rp64, row);
        end = gat_rp(p.row_ptr, p.rp64, row + 1);
        if (p.e_chunk > 0 && end - beg > p.e_chunk) return;  // done by the heavy items
    }

    float gi[NCH][VE];
    float c_part = 0.f;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const int dcol = (k * LANES + lig) * VE;
#pragma unroll
        for (int a = 0; a < VE; ++a) {
            float gv = 0.f, ov = 0.f;
            if (dcol + a < p.D) {
                gv = __ldg(p.g + row * p.ldg + head * p.D + dcol + a);
                ov = __ldg(p.out + row * p.ldo + head * p.D + dcol + a);
            }
            gi[k][a] = gv;
            c_part = fmaf(gv, ov, c_part);
        }
    }
    const float c_i = group_sum<LANES>(c_part, gmask);  // sum_k alpha_ik dalpha_ik = <g_i, out_i>
    const float el_i = __ldg(p.el + row * p.ld_e + head);
    const float m_i = __ldg(p.row_max + row * p.heads + head);
    const float l_i = __ldg(p.row_sum + row * p.heads + head);
    const float inv_l = l_i > 0.f ? 1.f / l_i : 0.f;
    float del_acc = 0.f;

    for (long long e0 = beg; e0 < end; e0 += LANES) {
        const int n = static_cast<int>(min(static_cast<long long>(LANES), end - e0));
        int my_c = 0;
        float my_alpha = 0.f, my_dact = 0.f;
        if (lig < n) {
            my_c = __ldg(p.col + e0 + lig);
            const float z = el_i + __ldg(p.er + static_cast<long long>(my_c) * p.ld_e + head);
            const float lz = z > 0.f ? z : p.slope * z;
            my_alpha = expf(p.sign * lz - m_i) * inv_l;
            my_dact = p.sign * (z > 0.f ? 1.f : p.slope);
        }
        float my_dalpha = 0.f;
        // kGatUnroll source rows in flight per lane before any reduction (the first version loaded one row, reduced,
        // then loaded the next: one dependent DRAM round trip per edge)
        for (int k0 = 0; k0 < n; k0 += kGatUnroll) {
            float part[kGatUnroll];
#pragma unroll
            for (int u = 0; u < kGatUnroll; ++u) {
                const long long c = __shfl_sync(gmask, my_c, (k0 + u) & (LANES - 1), LANES);
                float pr = 0.f;
                if (k0 + u < n) {
#pragma unroll
                    for (int k = 0; k < NCH; ++k) {
                        const int dcol = (k * LANES + lig) * VE;
                        if (dcol < p.D) {
                            const float* w = p.Wh + c * p.ldw + head * p.D + dcol;
                            if (VE == 4) {
                                const float4 v = ldg_nc_f4(w);
                                pr = fmaf(gi[k][0], v.x, pr);
                                pr = fmaf(gi[k][1 % VE], v.y, pr);
                                pr = fmaf(gi[k][2 % VE], v.z, pr);
                                pr = fmaf(gi[k][3 % VE], v.w, pr);
                            } else {
                                pr = fmaf(gi[k][0], ldg_nc_f1(w), pr);
                            }
                        }
                    }
                }
                part[u] = pr;
            }
#pragma unroll
            for (int u = 0; u < kGatUnroll; ++u) {
                const float dot = group_sum<LANES>(part[u], gmask);
                if (lig == k0 + u) my_dalpha = dot;
            }
        }
        if (lig < n) {
            // with dropout the aggregation weight is alpha * m' (m' = mask / (1-p)); c_i = <g_i, out_i> already
            // contains the same m' through the forward output
            const float mk = p.drop_thresh ? gat_keep_scale(p.seed, e0 + lig, head, p.heads, p.drop_thresh, p.drop_scale)
                                           : 1.f;
            const float dz = my_alpha * (mk * my_dalpha - c_i) * my_dact;
            p.ws[(e0 + lig) * p.heads + head] = make_float2(my_alpha * mk, dz);
            del_acc += dz;
        }
    }
    const float del = group_sum<LANES>(del_acc, gmask);
    if (lig == 0) {
        if (HEAVY) p.e_ws[unit * p.heads + head] = del;
        else p.d_el[row * p.ld_de + head] = del;
    }
}

// Pass 2: group = (src row j, head, slab) over the transposed CSR.
template <int VE, int LANES, bool HEAVY>
__global__ void __launch_bounds__(kGatThreads)
gat_backward_node_kernel(const GatBwdParams p, int n_slabs) {
    const int lig = threadIdx.x & (LANES - 1);
    const long long group = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) / LANES;
    const unsigned gmask = (LANES == 32) ? 0xffffffffu
                                         : (((1u << LANES) - 1u) << ((threadIdx.x & 31) & ~(LANES - 1)));
    const int per_row = p.heads * n_slabs;
    const long long unit = group / per_row;       // src row of the transposed CSR, or plan item when HEAVY
    const int rem = static_cast<int>(group - unit * per_row);
    const int head = rem / n_slabs;
    const int slab = rem - head * n_slabs;
    long long row, beg, end;
    if (HEAVY) {
        if (unit >= p.t_n_items) return;
        const int2 it = p.t_items[unit];
        row = it.x;
        const long long rb = gat_rp(p.t_row_ptr, p.rp64, row), re = gat_rp(p.t_row_ptr, p.rp64, row + 1);
        beg = rb + static_cast<long long>(it.y) * p.t_chunk;
        end = min(re, beg + p.t_chunk);
    } else {
        row = unit;
        if (row >= p.n_src) return;
        beg = gat_rp(p.t_row_ptr, p.rp64, row);
        end = gat_rp(p.t_row_ptr, p.rp64, row + 1);
        if (p.t_chunk > 0 && end - beg > p.t_chunk) return;  // done by the heavy items
    }
    const int dcol = (slab * LANES + lig) * VE;
    const bool lane_on = dcol < p.D;
    const float* __restrict__ gb = p.g + head * p.D + dcol;

    float acc[VE];
#pragma unroll
    for (int a = 0; a < VE; ++a) acc[a] = 0.f;
    float der_acc = 0.f;

    for (long long e0 = beg; e0 < end; e0 += LANES) {
        const int n = static_cast<int>(min(static_cast<long long>(LANES), end - e0));
        int my_i = 0;
        float my_alpha = 0.f;
        if (lig < n) {
            my_i = __ldg(p.t_col + e0 + lig);
            const long long eo = __ldg(p.perm + e0 + lig);
            const float2 ad = __ldg(p.ws + eo * p.heads + head);
            my_alpha = ad.x;
            der_acc += ad.y;
        }
        for (int k = 0; k < n; k += kGatUnroll) {
            float4 raw[kGatUnroll];
            float w[kGatUnroll];
#pragma unroll
            for (int u = 0; u < kGatUnroll; ++u) {
                const int src = (k + u) & (LANES - 1);
                const long long i = __shfl_sync(gmask, my_i, src, LANES);
                w[u] = __shfl_sync(gmask, my_alpha, src, LANES);
                if (k + u < n && lane_on) {
                    if (VE == 4) raw[u] = ldg_nc_f4(gb + i * p.ldg);
                    else raw[u].x = ldg_nc_f1(gb + i * p.ldg);
                } else {
                    raw[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    w[u] = 0.f;
                }
            }
#pragma unroll
            for (int u = 0; u < kGatUnroll; ++u) {
                acc[0] = fmaf(w[u], raw[u].x, acc[0]);
                if (VE == 4) {
                    acc[1] = fmaf(w[u], raw[u].y, acc[1]);
                    acc[2] = fmaf(w[u], raw[u].z, acc[2]);
                    acc[3] = fmaf(w[u], raw[u].w, acc[3]);
                }
            }
        }
    }
    const long long FDh = static_cast<long long>(p.heads) * p.D;
    if (slab == 0) {
        const float der = group_sum<LANES>(der_acc, gmask);
        if (lig == 0) {
            if (HEAVY) p.t_ws[unit * (FDh + p.heads) + FDh + head] = der;
            else p.d_er[row * p.ld_de + head] = der;
        }
    }
    if (!lane_on) return;
    float* __restrict__ o = HEAVY ? p.t_ws + unit * (FDh + p.heads) + head * p.D + dcol
                                  : p.d_Wh + row * p.ldd + head * p.D + dcol;
    const int valid = min(VE, p.D - dcol);
    if (VE == 4 && valid == 4 && !HEAVY) {
        stg_cs_f4(o, make_float4(acc[0], acc[1], acc[2], acc[3]));
    } else {
#pragma unroll
        for (int a = 0; a < VE; ++a)
            if (a < valid) o[a] = acc[a];
    }
}

// merge the partials of the long rows: one thread block per k == 0 item (its chunks are items [w, w + nc))
__global__ void gat_bwd_combine_edge_kernel(const GatBwdParams p) {
    const long long w = blockIdx.x;
    if (w >= p.e_n_items) return;
    const int2 it = p.e_items[w];
    if (it.y != 0) return;
    const long long row = it.x;
    const long long deg = gat_rp(p.row_ptr, p.rp64, row + 1) - gat_rp(p.row_ptr, p.rp64, row);
    const int nc = static_cast<int>((deg + p.e_chunk - 1) / p.e_chunk);
    for (int h = threadIdx.x; h < p.heads; h += blockDim.x) {
        float sacc = 0.f;
        for (int c = 0; c < nc; ++c) sacc += p.e_ws[(w + c) * p.heads + h];
        p.d_el[row * p.ld_de + h] = sacc;
    }
}
__global__ void gat_bwd_combine_node_kernel(const GatBwdParams p) {
    const long long w = blockIdx.x;
    if (w >= p.t_n_items) return;
    const int2 it = p.t_items[w];
    if (it.y != 0) return;
    const long long row = it.x;
    const long long deg = gat_rp(p.t_row_ptr, p.rp64, row + 1) - gat_rp(p.t_row_ptr, p.rp64, row);
    const int nc = static_cast<int>((deg + p.t_chunk - 1) / p.t_chunk);
    const long long FDh = static_cast<long long>(p.heads) * p.D, W = FDh + p.heads;
    for (long long f = threadIdx.x; f < W; f += blockDim.x) {
        float sacc = 0.f;
        for (int c = 0; c < nc; ++c) sacc += p.t_ws[(w + c) * W + f];
        if (f < FDh) p.d_Wh[row * p.ldd + f] = sacc;
        else p.d_er[row * p.ld_de + (f - FDh)] = sacc;
    }
}

// ------------------------------------------------- fused single-pass backward --
// The two-pass backward above streams BOTH the source rows Wh_j (pass 1, for dalpha_ij = <g_i, Wh_j>) and the gradient
// rows g_i (pass 2, for d_Wh_j = sum_i alpha_ij g_i) once per edge, and moves 8 bytes per (edge, head) through a
// workspace between the passes: products-shaped 55.5 ms against 19.7 ms for the forward.  Both per-edge quantities pair
// the SAME two rows, so one pass over the transposed CSR does it all: the warp that owns source row j keeps Wh_j and
// er_j in registers and streams the g_i rows of its edges ONCE — the dot product with the resident Wh_j gives dalpha_ij,
// the same registers feed d_Wh_j += alpha_ij g_i.  What the pass needs from destination i besides g_i is four scalars
// per head (el_i, row max, 1/row sum, c_i = <g_i, out_i>), packed by a node-level pre-pass into one 64-byte record that
// the scoring lane fetches with the column id.  dz_ij leaves as 4 bytes per (edge, head) at the edge's forward position
// (perm), d_er_j is its sum along the warp's own row, d_el_i its sum along forward row i (a 16-byte-per-edge streaming
// pass).  No atomics anywhere: every output element has one owner and a fixed summation order.
// Shapes: the whole-row ones of the forward (16-byte vectors, heads <= 4, 32 <= heads*D <= 512).

// sum four per-lane values over the warp with 6 shuffles: after the call every lane holds the total of value
// hq = 2*bit4(lane) + bit3(lane) (an 8-lane group per value)
__device__ __forceinline__ float warp_sum4(float v0, float v1, float v2, float v3, int lane) {
    const bool up16 = lane & 16;
    const float s0 = up16 ? v0 : v2, s1 = up16 ? v1 : v3;          // what the partner keeps
    float k0 = (up16 ? v2 : v0) + __shfl_xor_sync(0xffffffffu, s0, 16);
    float k1 = (up16 ? v3 : v1) + __shfl_xor_sync(0xffffffffu, s1, 16);
    const bool up8 = lane & 8;
    float r = (up8 ? k1 : k0) + __shfl_xor_sync(0xffffffffu, up8 ? k0 : k1, 8);
    r += __shfl_xor_sync(0xffffffffu, r, 4);
    r += __shfl_xor_sync(0xffffffffu, r, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    return r;
}

// node pre-pass: stats[i][h] = (el_i, row max, 1 / row sum, c_i = <g_i, out_i>) for h < heads, zeros above
__global__ void __launch_bounds__(128)
gat_bwd_prep_kernel(const GatBwdParams p, float4* __restrict__ stats) {
    const int lane = threadIdx.x & 31;
    const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (row >= p.n_dst) return;
    const int FD = p.heads * p.D;
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c0 = lane * 4; c0 < FD; c0 += 128) {
        const float4 gv = ldg_nc_f4(p.g + row * p.ldg + c0);
        const float4 ov = ldg_nc_f4(p.out + row * p.ldo + c0);
        const float d = gv.x * ov.x + gv.y * ov.y + gv.z * ov.z + gv.w * ov.w;
        const int h = c0 / p.D;
        c[0] += h == 0 ? d : 0.f; c[1] += h == 1 ? d : 0.f; c[2] += h == 2 ? d : 0.f; c[3] += h == 3 ? d : 0.f;
    }
    const float ci = warp_sum4(c[0], c[1], c[2], c[3], lane);
    if ((lane & 7) == 0) {
        const int hq = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
        float4 st = make_float4(0.f, 0.f, 0.f, 0.f);
        if (hq < p.heads) {
            const float l = __ldg(p.row_sum + row * p.heads + hq);
            st = make_float4(__ldg(p.el + row * p.ld_e + hq), __ldg(p.row_max + row * p.heads + hq),
                             l > 0.f ? 1.f / l : 0.f, ci);
        }
        stats[row * 4 + hq] = st;
    }
}

// HEAVY = false: warp w owns row w of the transposed CSR (source node j); rows longer than p.t_chunk are skipped when a
//                plan is given.   HEAVY = true: warp w owns plan item w = (row, k) and writes partial d_Wh / d_er rows
//                to p.t_ws, merged by gat_bwd_combine_node_kernel.
// P = g rows in flight per lane (rolling: the slot of a consumed edge is refilled with edge + P of the same 32-edge
// chunk).  First version (profiles/r02_gat_bwd_fused.txt): load 4 rows, reduce, load 4 more — ~110 warp instructions per
// edge (per-edge head selects, 64-bit row address multiplies, zero fills), issue active 35 %, DRAM 42-55 %, every top
// stall on the first use of a loaded row.  Now: head masks are hoisted out of the edge loop (FFMA instead of
// ISETP+FSEL), the scoring lane computes the row's byte offset once, the chunk's first P rows are requested BEFORE the
// scores are computed, and nothing is zero-filled (short groups are skipped by warp-uniform branches).
template <int NV, bool HEAVY, int P>
__device__ __forceinline__ void gat_backward_fused_body(const GatBwdParams& p, const float4* __restrict__ stats,
                                                        float* __restrict__ dz_ws, const long long wid, long long* s_off,
                                                        float (*s_a)[4], float2 (*s_bc)[4], float (*s_dz)[4]) {
    static_assert(32 % P == 0, "window depth must divide the 32-edge chunk");
    const int lane = threadIdx.x;
    const int H = p.heads, FD = p.heads * p.D;
    long long row, beg, end;
    if (HEAVY) {
        const int2 it = p.t_items[wid];
        row = it.x;
        const long long rb = gat_rp(p.t_row_ptr, p.rp64, row), re = gat_rp(p.t_row_ptr, p.rp64, row + 1);
        beg = rb + static_cast<long long>(it.y) * p.t_chunk;
        end = min(re, beg + p.t_chunk);
    } else {
        row = wid;
        if (row >= p.n_src) return;
        beg = gat_rp(p.t_row_ptr, p.rp64, row);
        end = gat_rp(p.t_row_ptr, p.rp64, row + 1);
        if (p.t_chunk > 0 && end - beg > p.t_chunk) return;  // done by the heavy items
    }
    const int hq = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);   // the head whose reductions land in this lane

    int hsel[NV];
    bool von[NV];
    float4 wh[NV];
    float hm[NV][4];                                            // 1 where vector i belongs to head h
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int c0 = (i * 32 + lane) * 4;
        von[i] = c0 < FD;
        hsel[i] = von[i] ? c0 / p.D : 0;
        wh[i] = von[i] ? ldg_nc_f4(p.Wh + row * p.ldw + c0) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int h = 0; h < 4; ++h) hm[i][h] = (von[i] && hsel[i] == h) ? 1.f : 0.f;
    }
    float er_j[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) er_j[h] = h < H ? __ldg(p.er + row * p.ld_e + h) : 0.f;

    float acc[NV][4];
    float4 raw[P][NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
#pragma unroll
        for (int a = 0; a < 4; ++a) acc[i][a] = 0.f;
#pragma unroll
        for (int u = 0; u < P; ++u) raw[u][i] = make_float4(0.f, 0.f, 0.f, 0.f);   // vectors past FD stay zero
    }
    float der = 0.f;
    const char* __restrict__ gbase = reinterpret_cast<const char*>(p.g) + lane * 16;
    const long long g_stride = p.ldg * 4;

    // per-lane edge of the current chunk: destination id, forward position, destination record; ids two chunks ahead
    int i_cur = 0, eo_cur = 0, i_nxt = 0, eo_nxt = 0;
    float4 st[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) st[h] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (beg + lane < end) {
        i_cur = __ldg(p.t_col + beg + lane);
        eo_cur = __ldg(p.perm + beg + lane);
#pragma unroll
        for (int h = 0; h < 4; ++h) st[h] = __ldg(stats + static_cast<long long>(i_cur) * 4 + h);
    }
    if (beg + 32 + lane < end) {
        i_nxt = __ldg(p.t_col + beg + 32 + lane);
        eo_nxt = __ldg(p.perm + beg + 32 + lane);
    }

    for (long long e0 = beg; e0 < end; e0 += 32) {
        const int n = static_cast<int>(min(32ll, end - e0));
        s_off[lane] = static_cast<long long>(i_cur) * g_stride;
        __syncwarp();
        // the chunk's first P gradient rows leave now; the scores are computed under their latency
#pragma unroll
        for (int u = 0; u < P; ++u) {
            if (u < n) {
                const char* src = gbase + s_off[u];
#pragma unroll
                for (int i = 0; i < NV; ++i)
                    if (von[i]) raw[u][i] = ldg_nc_f4(reinterpret_cast<const float*>(src + i * 512));
            }
        }
        // score this lane's edge for every head: aggregation weight a = alpha*mask, dz = b*dalpha - cc
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            float a = 0.f, b = 0.f, cc = 0.f;
            if (h < H && lane < n) {
                const float z = st[h].x + er_j[h];
                const float lz = z > 0.f ? z : p.slope * z;
                const float alpha = expf(p.sign * lz - st[h].y) * st[h].z;
                const float dact = p.sign * (z > 0.f ? 1.f : p.slope);
                const float mk = p.drop_thresh ? gat_keep_scale(p.seed, eo_cur, h, H, p.drop_thresh, p.drop_scale) : 1.f;
                a = alpha * mk;
                b = a * dact;
                cc = alpha * dact * st[h].w;
            }
            s_a[lane][h] = a;
            s_bc[lane][h] = make_float2(b, cc);
        }
        const int eo_mine = eo_cur;
        __syncwarp();
        // next chunk's records start their round trip now; its ids were fetched a chunk ago
        i_cur = i_nxt;
        eo_cur = eo_nxt;
        if (e0 + 32 + lane < end) {
#pragma unroll
            for (int h = 0; h < 4; ++h) st[h] = __ldg(stats + static_cast<long long>(i_cur) * 4 + h);
        }
        if (e0 + 64 + lane < end) {
            i_nxt = __ldg(p.t_col + e0 + 64 + lane);
            eo_nxt = __ldg(p.perm + e0 + 64 + lane);
        }

#pragma unroll 1
        for (int k = 0; k < n; k += P) {
#pragma unroll
            for (int u = 0; u < P; ++u) {
                const int e = k + u;
                if (e < n) {
                    float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
#pragma unroll
                    for (int i = 0; i < NV; ++i) {
                        const float d = raw[u][i].x * wh[i].x + raw[u][i].y * wh[i].y + raw[u][i].z * wh[i].z +
                                        raw[u][i].w * wh[i].w;
                        v0 = fmaf(hm[i][0], d, v0); v1 = fmaf(hm[i][1], d, v1);
                        v2 = fmaf(hm[i][2], d, v2); v3 = fmaf(hm[i][3], d, v3);
                    }
                    const float dalpha = warp_sum4(v0, v1, v2, v3, lane);
                    const float2 bc = s_bc[e][hq];
                    const float dz = fmaf(bc.x, dalpha, -bc.y);
                    der += dz;
                    if ((lane & 7) == 0) s_dz[e][hq] = dz;
#pragma unroll
                    for (int i = 0; i < NV; ++i) {
                        const float a = s_a[e][hsel[i]];
                        acc[i][0] = fmaf(a, raw[u][i].x, acc[i][0]);
                        acc[i][1] = fmaf(a, raw[u][i].y, acc[i][1]);
                        acc[i][2] = fmaf(a, raw[u][i].z, acc[i][2]);
                        acc[i][3] = fmaf(a, raw[u][i].w, acc[i][3]);
                    }
                    if (e + P < n) {                             // refill the slot with edge e + P of this chunk
                        const char* src = gbase + s_off[e + P];
#pragma unroll
                        for (int i = 0; i < NV; ++i)
                            if (von[i]) raw[u][i] = ldg_nc_f4(reinterpret_cast<const float*>(src + i * 512));
                    }
                }
            }
        }
        __syncwarp();
        if (lane < n) {   // dz of this lane's edge, all heads, at the edge's position in the forward CSR
            if (H == 4) {
                *reinterpret_cast<float4*>(dz_ws + static_cast<long long>(eo_mine) * 4) =
                    *reinterpret_cast<const float4*>(&s_dz[lane][0]);
            } else {
                for (int h = 0; h < H; ++h) dz_ws[static_cast<long long>(eo_mine) * H + h] = s_dz[lane][h];
            }
        }
        __syncwarp();  // s_* are rewritten by the next chunk
    }

    const long long FDh = FD;
    if ((lane & 7) == 0 && hq < H) {
        if (HEAVY) p.t_ws[wid * (FDh + H) + FDh + hq] = der;
        else p.d_er[row * p.ld_de + hq] = der;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        if (!von[i]) continue;
        const int c0 = (i * 32 + lane) * 4;
        if (HEAVY) {
            float* o = p.t_ws + wid * (FDh + H) + c0;
            o[0] = acc[i][0]; o[1] = acc[i][1]; o[2] = acc[i][2]; o[3] = acc[i][3];
        } else {
            stg_cs_f4(p.d_Wh + row * p.ldd + c0, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
        }
    }
}

// ONE launch: blocks [0, t_n_items) take the plan items of the long rows (they start first), the rest one row each — the
// short rows' dependent start-up round trips then overlap the long rows' streaming instead of following it (as two
// launches: 14.8 ms at 60 % DRAM, then 18.2 ms at 39 %; profiles/r02_gat_bwd_fused_v2.txt).  Both bodies stay
// compile-time specialised.
template <int NV, int P>
__global__ void __launch_bounds__(32, 16)
gat_backward_fused_kernel(const GatBwdParams p, const float4* __restrict__ stats, float* __restrict__ dz_ws) {
    __shared__ long long s_off[32];
    __shared__ float s_a[32][4];
    __shared__ float2 s_bc[32][4];
    __shared__ __align__(16) float s_dz[32][4];
    if (static_cast<long long>(blockIdx.x) < p.t_n_items)
        gat_backward_fused_body<NV, true, P>(p, stats, dz_ws, blockIdx.x, s_off, s_a, s_bc, s_dz);
    else
        gat_backward_fused_body<NV, false, P>(p, stats, dz_ws, static_cast<long long>(blockIdx.x) - p.t_n_items, s_off, s_a, s_bc, s_dz);
}

// d_el_i = sum of dz over forward row i (dz_ws is [nnz, heads], CSR edge order): warp per row / per plan item
template <bool HEAVY>
__global__ void __launch_bounds__(128)
gat_bwd_del_kernel(const GatBwdParams p, const float* __restrict__ dz_ws) {
    const int lane = threadIdx.x & 31;
    const long long unit = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int H = p.heads;
    long long row, beg, end;
    if (HEAVY) {
        if (unit >= p.e_n_items) return;
        const int2 it = p.e_items[unit];
        row = it.x;
        const long long rb = gat_rp(p.row_ptr, p.rp64, row), re = gat_rp(p.row_ptr, p.rp64, row + 1);
        beg = rb + static_cast<long long>(it.y) * p.e_chunk;
        end = min(re, beg + p.e_chunk);
    } else {
        row = unit;
        if (row >= p.n_dst) return;
        beg = gat_rp(p.row_ptr, p.rp64, row);
        end = gat_rp(p.row_ptr, p.rp64, row + 1);
        if (p.e_chunk > 0 && end - beg > p.e_chunk) return;
    }
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    if (H == 4) {
        const float4* __restrict__ w4 = reinterpret_cast<const float4*>(dz_ws);
        long long e = beg + lane;
        for (; e + 96 < end; e += 128) {
            const float4 a = __ldg(w4 + e), b = __ldg(w4 + e + 32), c = __ldg(w4 + e + 64), d = __ldg(w4 + e + 96);
            s[0] += (a.x + b.x) + (c.x + d.x); s[1] += (a.y + b.y) + (c.y + d.y);
            s[2] += (a.z + b.z) + (c.z + d.z); s[3] += (a.w + b.w) + (c.w + d.w);
        }
        for (; e < end; e += 32) {
            const float4 a = __ldg(w4 + e);
            s[0] += a.x; s[1] += a.y; s[2] += a.z; s[3] += a.w;
        }
    } else {
        for (long long e = beg + lane; e < end; e += 32)
            for (int h = 0; h < H; ++h) s[h] += __ldg(dz_ws + e * H + h);
    }
    const float t = warp_sum4(s[0], s[1], s[2], s[3], lane);
    const int hq = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
    if ((lane & 7) == 0 && hq < H) {
        if (HEAVY) p.e_ws[unit * H + hq] = t;
        else p.d_el[row * p.ld_de + hq] = t;
    }
}

template <int NV, int P>
static int launch_gat_bwd_fused(const GatBwdParams& p, cudaStream_t st) {
    float4* stats = nullptr;
    DGLLB_CUDA_TRY(cudaMallocAsync(&stats, sizeof(float4) * 4 * static_cast<size_t>(p.n_dst > 0 ? p.n_dst : 1), st));
    float* dz_ws = reinterpret_cast<float*>(p.ws);
    int rc = DGLLB_OK;
    do {
        if (p.n_dst > 0) {
            const long long blocks = (p.n_dst * 32 + 127) / 128;
            if (blocks >= (1ll << 31)) { set_error("gat_backward: grid too large"); rc = DGLLB_ERR_INVALID; break; }
            gat_bwd_prep_kernel<<<static_cast<unsigned>(blocks), 128, 0, st>>>(p, stats);
        }
        if (p.n_src + p.t_n_items >= (1ll << 31)) { set_error("gat_backward: grid too large"); rc = DGLLB_ERR_INVALID; break; }
        if (p.n_src + p.t_n_items > 0)
            gat_backward_fused_kernel<NV, P><<<static_cast<unsigned>(p.n_src + p.t_n_items), 32, 0, st>>>(p, stats, dz_ws);
        if (p.t_n_items > 0) gat_bwd_combine_node_kernel<<<static_cast<unsigned>(p.t_n_items), 256, 0, st>>>(p);
        if (p.e_n_items > 0)
            gat_bwd_del_kernel<true><<<static_cast<unsigned>((p.e_n_items * 32 + 127) / 128), 128, 0, st>>>(p, dz_ws);
        if (p.n_dst > 0)
            gat_bwd_del_kernel<false><<<static_cast<unsigned>((p.n_dst * 32 + 127) / 128), 128, 0, st>>>(p, dz_ws);
        if (p.e_n_items > 0) gat_bwd_combine_edge_kernel<<<static_cast<unsigned>(p.e_n_items), 32, 0, st>>>(p);
        g_launch_count.fetch_add((p.n_dst > 0 ? 2 : 0) + (p.n_src + p.t_n_items > 0 ? 1 : 0) + (p.t_n_items > 0 ? 1 : 0) +
                                 (p.e_n_items > 0 ? 2 : 0));
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { set_error("gat_backward: %s", cudaGetErrorString(e)); rc = DGLLB_ERR_CUDA; }
    } while (0);
    cudaFreeAsync(stats, st);
    return rc;
}

static int gat_bwd_block_threads() {
    const int tb = opt_get(OPT_GAT_BWD_TB);  // default 64: small blocks retire evenly on ragged rows, 64.0 -> 55.9 ms (products-shaped)
    return (tb == 32 || tb == 64 || tb == 128 || tb == 256) ? tb : 64;
}

template <int VE, int LANES, bool HEAVY>
static int launch_gat_bwd_edge(const GatBwdParams& p, int nch, cudaStream_t st) {
    const long long groups = (HEAVY ? p.e_n_items : p.n_dst) * p.heads;
    if (groups == 0) return DGLLB_OK;
    const int tb = gat_bwd_block_threads();
    const int gpb = tb / LANES;
    const long long blocks = (groups + gpb - 1) / gpb;
    DGLLB_REQUIRE(blocks < (1ll << 31), "gat_backward: grid too large");
    const unsigned g = static_cast<unsigned>(blocks);
    if (nch <= 1) gat_backward_edge_kernel<VE, LANES, 1, HEAVY><<<g, tb, 0, st>>>(p);
    else if (nch <= 2) gat_backward_edge_kernel<VE, LANES, 2, HEAVY><<<g, tb, 0, st>>>(p);
    else if (nch <= 4) gat_backward_edge_kernel<VE, LANES, 4, HEAVY><<<g, tb, 0, st>>>(p);
    else {
        set_error("gat_backward: head width D=%d too large for this build", p.D);
        return DGLLB_ERR_UNSUPPORTED;
    }
    DGLLB_LAUNCH_CHECK();
    return DGLLB_OK;
}

template <int VE, int LANES, bool HEAVY>
static int launch_gat_bwd_node(const GatBwdParams& p, int nch, cudaStream_t st) {
    const long long groups = (HEAVY ? p.t_n_items : p.n_src) * p.heads * nch;
    if (groups == 0) return DGLLB_OK;
    const int tb = gat_bwd_block_threads();
    const int gpb = tb / LANES;
    const long long blocks = (groups + gpb - 1) / gpb;
    DGLLB_REQUIRE(blocks < (1ll << 31), "gat_backward: grid too large");
    gat_backward_node_kernel<VE, LANES, HEAVY><<<static_cast<unsigned>(blocks), tb, 0, st>>>(p, nch);
    DGLLB_LAUNCH_CHECK();
    return DGLLB_OK;
}

template <int VE, int LANES>
static int launch_gat_bwd_lanes(const GatBwdParams& p, int nch, cudaStream_t st) {
    int rc;
    if (p.e_n_items > 0) {
        if ((rc = launch_gat_bwd_edge<VE, LANES, true>(p, nch, st)) != DGLLB_OK) return rc;
    }
    if ((rc = launch_gat_bwd_edge<VE, LANES, false>(p, nch, st)) != DGLLB_OK) return rc;
    if (p.e_n_items > 0) {
        gat_bwd_combine_edge_kernel<<<static_cast<unsigned>(p.e_n_items), 32, 0, st>>>(p);
        DGLLB_LAUNCH_CHECK();
    }
    if (p.t_n_items > 0) {
        if ((rc = launch_gat_bwd_node<VE, LANES, true>(p, nch, st)) != DGLLB_OK) return rc;
    }
    if ((rc = launch_gat_bwd_node<VE, LANES, false>(p, nch, st)) != DGLLB_OK) return rc;
    if (p.t_n_items > 0) {
        gat_bwd_combine_node_kernel<<<static_cast<unsigned>(p.t_n_items), 256, 0, st>>>(p);
        DGLLB_LAUNCH_CHECK();
    }
    return DGLLB_OK;
}

template <int VE>
static int launch_gat_bwd(GatBwdParams& p, const dgllb_csr_plan* plan, const dgllb_csr_plan* t_plan, cudaStream_t st) {
    int lanes = 32;
    if (VE > 1 && p.D <= 8 * VE) lanes = 8;
    else if (VE > 1 && p.D <= 16 * VE) lanes = 16;
    const int nch = (p.D + lanes * VE - 1) / (lanes * VE);
    const bool eh = plan && plan->n_heavy_rows > 0, th = t_plan && t_plan->n_heavy_rows > 0;
    p.e_chunk = eh ? plan->chunk_edges : 0;
    p.e_items = eh ? plan->items : nullptr;
    p.e_n_items = eh ? plan->n_items : 0;
    p.t_chunk = th ? t_plan->chunk_edges : 0;
    p.t_items = th ? t_plan->items : nullptr;
    p.t_n_items = th ? t_plan->n_items : 0;
    p.e_ws = p.t_ws = nullptr;
    char* ws = nullptr;
    const size_t FDh = static_cast<size_t>(p.heads) * p.D;
    const size_t b_e = (sizeof(float) * static_cast<size_t>(p.e_n_items) * p.heads + 255) & ~static_cast<size_t>(255);
    const size_t b_t = sizeof(float) * static_cast<size_t>(p.t_n_items) * (FDh + p.heads);
    if (eh || th) {
        DevInfo di_;
        int rc_ = get_devinfo(&di_);
        if (rc_ != DGLLB_OK) return rc_;
        DGLLB_CUDA_TRY(cudaMallocAsync(&ws, b_e + b_t + 256, st));
        p.e_ws = reinterpret_cast<float*>(ws);
        p.t_ws = reinterpret_cast<float*>(ws + b_e);
    }
    int rc;
    // fused single pass over CSR^T for the whole-row shapes (option gat_bwd_kernel=twopass pins the older path)
    const int FD = p.heads * p.D;
    if (VE == 4 && p.heads <= 4 && FD >= 32 && FD <= 512 && opt_get(OPT_GAT_BWD_KERNEL) != 1) {
        const int depth = opt_get(OPT_GAT_BWD_DEPTH);
        if (FD <= 128) rc = depth == 4 ? launch_gat_bwd_fused<1, 4>(p, st) : launch_gat_bwd_fused<1, 8>(p, st);
        else if (FD <= 256) rc = depth == 4 ? launch_gat_bwd_fused<2, 4>(p, st) : launch_gat_bwd_fused<2, 8>(p, st);
        else rc = depth == 2 ? launch_gat_bwd_fused<4, 2>(p, st) : launch_gat_bwd_fused<4, 4>(p, st);
    } else
    if (lanes == 8) rc = launch_gat_bwd_lanes<VE, 8>(p, nch, st);
    else if (lanes == 16) rc = launch_gat_bwd_lanes<VE, 16>(p, nch, st);
    else rc = launch_gat_bwd_lanes<VE, 32>(p, nch, st);
    if (ws) cudaFreeAsync(ws, st);
    return rc;
}

template <int VE>
static int launch_gat_fwd(GatParams& p, const dgllb_csr_plan* plan, cudaStream_t st) {
    // whole-row kernel: all heads of a row in one warp (vector path, <= 4 heads, 32 <= heads*D <= 512; the only forward
    // kernel that takes the long-row plan, which is what a 47-class output layer on a skewed graph needs).  On a uniform
    // degree-50 graph it runs at the SpMM rate (20.5 ms vs 19.2 ms, products-sized); on skewed graphs it needs the
    // nnz-split plan for long rows (a 89K-edge row would otherwise be one warp's job).  DGLLB_GAT_KERNEL=group|row pins.
    const int FD = p.heads * p.D;
    const int force = opt_get(OPT_GAT_KERNEL);   // 1 = generic lane-group kernel, 2 = whole-row kernel
    const bool row_ok = VE == 4 && p.heads <= 4 && FD >= 32 && FD <= 512;  // below 32 floats most lanes of a warp would idle
    if (row_ok && force != 1) {
        const bool heavy = plan && plan->n_heavy_rows > 0;
        float* ws = nullptr;
        p.chunk = heavy ? plan->chunk_edges : 0;
        p.items = heavy ? plan->items : nullptr;
        p.n_items = heavy ? plan->n_items : 0;
        if (heavy) {
            DevInfo di_;
            int rc_ = get_devinfo(&di_);
            if (rc_ != DGLLB_OK) return rc_;
            DGLLB_CUDA_TRY(cudaMallocAsync(&ws, sizeof(float) * static_cast<size_t>(p.n_items) * (FD + 8), st));
        }
        p.ws = ws;
        // one warp per block: with 8 warps a block kept its registers until its longest row was done (achieved
        // occupancy 27 %, profiles/r01_gat_fwd_bwd.txt); 4 edges in flight per warp (8 spills)
        int cw = opt_get(OPT_GAT_ROW_WARPS);  // option gat_row_warps=1|4|8 pins the block size (measurement aid)
        if (cw <= 0) cw = 1;
        const long long blocks = (p.n_dst + cw - 1) / cw;
        const long long hblocks = (p.n_items + cw - 1) / cw;
        DGLLB_REQUIRE(blocks < (1ll << 31) && hblocks < (1ll << 31), "gat_forward: grid too large");
        const unsigned g = static_cast<unsigned>(blocks), hg = static_cast<unsigned>(hblocks);
#define DGLLB_GAT_ROW3(NVV, UU, WW)                                                               \
        do {                                                                                      \
            if (heavy) gat_forward_row_kernel<NVV, true, UU, WW><<<hg, WW * 32, 0, st>>>(p);      \
            gat_forward_row_kernel<NVV, false, UU, WW><<<g, WW * 32, 0, st>>>(p);                 \
        } while (0)
#define DGLLB_GAT_ROW(NVV)                                                                        \
        do {                                                                                      \
            if (cw == 8) DGLLB_GAT_ROW3(NVV, 4, 8);                                               \
            else if (cw == 4) DGLLB_GAT_ROW3(NVV, 4, 4);                                          \
            else DGLLB_GAT_ROW3(NVV, 4, 1);                                                       \
        } while (0)
        if (FD <= 128) DGLLB_GAT_ROW(1);
        else if (FD <= 256) DGLLB_GAT_ROW(2);
        else DGLLB_GAT_ROW(4);
#undef DGLLB_GAT_ROW3
#undef DGLLB_GAT_ROW
        g_launch_count.fetch_add(heavy ? 1 : 0);
        DGLLB_LAUNCH_CHECK();
        if (heavy) {
            gat_combine_heavy_kernel<<<static_cast<unsigned>((p.n_items * 32 + 255) / 256), 256, 0, st>>>(p);
            g_launch_count.fetch_add(1);
            cudaError_t e = cudaGetLastError();
            cudaFreeAsync(ws, st);
            if (e != cudaSuccess) { set_error("gat_forward: %s", cudaGetErrorString(e)); return DGLLB_ERR_CUDA; }
        }
        return DGLLB_OK;
    }
    int lanes = 32;
    if (VE > 1 && p.D <= 8 * VE) lanes = 8;
    else if (VE > 1 && p.D <= 16 * VE) lanes = 16;
    p.n_slabs = (p.D + lanes * VE - 1) / (lanes * VE);
    const long long groups = p.n_dst * p.heads * p.n_slabs;
    const int gpb = kGatThreads / lanes;
    const long long blocks = (groups + gpb - 1) / gpb;
    DGLLB_REQUIRE(blocks < (1ll << 31), "gat_forward: grid too large");
    const unsigned g = static_cast<unsigned>(blocks);
    if (lanes == 8) gat_forward_kernel<VE, 8><<<g, kGatThreads, 0, st>>>(p);
    else if (lanes == 16) gat_forward_kernel<VE, 16><<<g, kGatThreads, 0, st>>>(p);
    else gat_forward_kernel<VE, 32><<<g, kGatThreads, 0, st>>>(p);
    DGLLB_LAUNCH_CHECK();
    return DGLLB_OK;
}

__global__ void gat_dropout_mask_kernel(unsigned long long seed, long long nnz, int H, unsigned thresh, float scale,
                                        float* __restrict__ out) {
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nnz * H) return;
    out[i] = gat_keep_scale(seed, i / H, static_cast<int>(i % H), H, thresh, scale);
}

}  // namespace dgllb

using namespace dgllb;

extern "C" int dgllb_gat_dropout_mask(uint64_t drop_seed, int64_t nnz, int heads, float drop_p, float* mask_out,
                                      void* stream) {
    DGLLB_REQUIRE(nnz >= 0 && heads >= 1 && drop_p >= 0.f && drop_p < 1.f && (nnz == 0 || mask_out),
                  "gat_dropout_mask: bad arguments");
    if (nnz == 0) return DGLLB_OK;
    unsigned th = drop_p > 0.f ? static_cast<unsigned>(static_cast<double>(drop_p) * 4294967296.0) : 0u;
    if (drop_p > 0.f && th == 0u) th = 1u;
    const long long total = nnz * heads;
    gat_dropout_mask_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        drop_seed, nnz, heads, th, 1.f / (1.f - drop_p), mask_out);
    DGLLB_LAUNCH_CHECK();
    return DGLLB_OK;
}

extern "C" int dgllb_gat_forward(const void* row_ptr, int row_ptr_is64, const int32_t* col_idx,
                                 const float* Wh, int64_t ldw, const float* el, const float* er,
                                 int64_t ld_e, float* out, int64_t ldo, float* row_max, float* row_sum,
                                 int64_t n_dst, int64_t n_src, int heads, int D, float slope,
                                 int mode, int epilogue, float drop_p, uint64_t drop_seed,
                                 const dgllb_csr_plan* plan, void* stream) {
    DGLLB_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "gat_forward: dropout probability must be in [0, 1)");
    DGLLB_REQUIRE(n_dst >= 0 && n_src >= 0 && heads >= 1 && D >= 1, "gat_forward: bad sizes");
    DGLLB_REQUIRE(!plan || plan->n_rows == n_dst, "gat_forward: plan was built for another row count");
    if (n_dst == 0) return DGLLB_OK;
    DGLLB_REQUIRE(row_ptr && Wh && el && er && out, "gat_forward: null pointer");
    DGLLB_REQUIRE(ldw >= static_cast<int64_t>(heads) * D && ldo >= static_cast<int64_t>(heads) * D,
                  "gat_forward: leading dimension smaller than heads*D");
    DGLLB_REQUIRE(ld_e >= heads, "gat_forward: ld_e smaller than heads");
    DGLLB_REQUIRE(mode == DGLLB_GAT_SOFTMAX || mode == DGLLB_GAT_EXP_NEG, "gat_forward: unknown mode %d", mode);
    GatParams p;
    p.row_ptr = row_ptr; p.rp64 = row_ptr_is64; p.col = col_idx; p.Wh = Wh; p.ldw = ldw;
    p.el = el; p.er = er; p.ld_e = ld_e; p.out = out; p.ldo = ldo; p.row_max = row_max; p.row_sum = row_sum;
    p.n_dst = n_dst; p.heads = heads; p.D = D; p.n_slabs = 1; p.slope = slope;
    p.sign = mode == DGLLB_GAT_SOFTMAX ? 1.f : -1.f;
    p.epi = epilogue;
    p.drop_thresh = drop_p > 0.f ? static_cast<unsigned>(static_cast<double>(drop_p) * 4294967296.0) : 0u;
    if (drop_p > 0.f && p.drop_thresh == 0u) p.drop_thresh = 1u;
    p.drop_scale = 1.f / (1.f - drop_p);
    p.seed = drop_seed;
    p.chunk = 0; p.items = nullptr; p.n_items = 0; p.ws = nullptr;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec = aligned16(Wh) && aligned16(out) && ldw % 4 == 0 && ldo % 4 == 0 && D % 4 == 0;
    return vec ? launch_gat_fwd<4>(p, plan, st) : launch_gat_fwd<1>(p, plan, st);
}

extern "C" int dgllb_gat_backward(const void* row_ptr, int row_ptr_is64, const int32_t* col_idx,
                                  const void* t_row_ptr, const int32_t* t_col_idx, const int32_t* perm,
                                  const float* Wh, int64_t ldw, const float* el, const float* er,
                                  int64_t ld_e, const float* out, int64_t ldo, const float* row_max,
                                  const float* row_sum, const float* g, int64_t ldg, float* d_Wh,
                                  int64_t ldd, float* d_el, float* d_er, int64_t ld_de, float* edge_ws,
                                  int64_t n_dst, int64_t n_src, int heads, int D, float slope,
                                  int mode, float drop_p, uint64_t drop_seed, const dgllb_csr_plan* plan,
                                  const dgllb_csr_plan* t_plan, void* stream) {
    DGLLB_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "gat_backward: dropout probability must be in [0, 1)");
    DGLLB_REQUIRE((!plan || plan->n_rows == n_dst) && (!t_plan || t_plan->n_rows == n_src),
                  "gat_backward: plan built for another row count");
    DGLLB_REQUIRE(n_dst >= 0 && n_src >= 0 && heads >= 1 && D >= 1, "gat_backward: bad sizes");
    if (n_dst == 0 && n_src == 0) return DGLLB_OK;
    DGLLB_REQUIRE(row_ptr && t_row_ptr && Wh && el && er && out && row_max && row_sum && g && d_Wh && d_el &&
                      d_er && edge_ws,
                  "gat_backward: null pointer");
    DGLLB_REQUIRE(mode == DGLLB_GAT_SOFTMAX || mode == DGLLB_GAT_EXP_NEG, "gat_backward: unknown mode %d", mode);
    const int64_t FD = static_cast<int64_t>(heads) * D;
    DGLLB_REQUIRE(ldw >= FD && ldo >= FD && ldg >= FD && ldd >= FD && ld_e >= heads && ld_de >= heads,
                  "gat_backward: leading dimension too small");
    GatBwdParams p;
    p.row_ptr = row_ptr; p.rp64 = row_ptr_is64; p.col = col_idx; p.t_row_ptr = t_row_ptr; p.t_col = t_col_idx;
    p.perm = perm; p.Wh = Wh; p.ldw = ldw; p.el = el; p.er = er; p.ld_e = ld_e; p.out = out; p.ldo = ldo;
    p.row_max = row_max; p.row_sum = row_sum; p.g = g; p.ldg = ldg; p.d_Wh = d_Wh; p.ldd = ldd;
    p.d_el = d_el; p.d_er = d_er; p.ld_de = ld_de; p.ws = reinterpret_cast<float2*>(edge_ws);
    p.n_dst = n_dst; p.n_src = n_src; p.heads = heads; p.D = D; p.slope = slope;
    p.sign = mode == DGLLB_GAT_SOFTMAX ? 1.f : -1.f;
    p.drop_thresh = drop_p > 0.f ? static_cast<unsigned>(static_cast<double>(drop_p) * 4294967296.0) : 0u;
    if (drop_p > 0.f && p.drop_thresh == 0u) p.drop_thresh = 1u;
    p.drop_scale = 1.f / (1.f - drop_p);
    p.seed = drop_seed;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec = aligned16(Wh) && aligned16(g) && aligned16(d_Wh) && ldw % 4 == 0 && ldg % 4 == 0 &&
                     ldd % 4 == 0 && D % 4 == 0 && (reinterpret_cast<uintptr_t>(edge_ws) & 7) == 0;
    return vec ? launch_gat_bwd<4>(p, plan, t_plan, st) : launch_gat_bwd<1>(p, plan, t_plan, st);
}
