// binspmm.cu — bit-packed binarized neighbourhood aggregation (sm_100a).
//
// The reference only names this feature (README.md:11, "Quantization/Bineraztion"
// box in DGLL Architecture.drawio); semantics are fixed in SURVEY.md §8 a18:
//   X_b = (X >= 0) packed along the feature axis, 32 features per uint32 word;
//   cnt[i,f] = sum_{j in N(i)} X_b[j,f]           (int32, exact)
//   sum_pm1 = 2*cnt - deg_i ; mean_pm1 = sum_pm1 / deg_i  (fp32 epilogue)
//
// Formulation: one warp per destination row, lane l owns packed word l of every neighbour; per-feature counts are
// kept bit-sliced in carry-save form (see the kernel comment) — the per-feature popcount over neighbours is computed
// 32 features at a time with bitwise adders, then scattered feature-major with shuffles.  HBM traffic per edge is
// 4 B (index) + 4*words B (packed row): 76..80 B instead of 2,408 B at F=602; the packed Reddit-shaped table
// (18.6 MB) is L2 resident, so the kernel is bound by instruction issue / L2, not by HBM.
// Algorithmic bytes: nnz*(4 + 4*ceil(F/32)) + n_dst*(4*F + r).
#include "common.cuh"
#include <stdlib.h>
#include "internal.cuh"

namespace dgllb {

__device__ __forceinline__ long long bin_rp(const void* p, int is64, long long i) {
    return is64 ? reinterpret_cast<const long long*>(p)[i]
                : static_cast<long long>(reinterpret_cast<const int*>(p)[i]);
}

// (row, word) items; lanes read 32 consecutive floats, ballot packs them.
__global__ void __launch_bounds__(256)
binarize_pack_kernel(const float* __restrict__ X, long long ldx, uint32_t* __restrict__ packed,
                     long long wpr, long long n_rows, int F) {
    const int lane = threadIdx.x & 31;
    const long long item = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (item >= n_rows * wpr) return;
    const long long r = item / wpr;
    const int w = static_cast<int>(item - r * wpr);
    const int f = w * 32 + lane;
    bool bit = false;
    if (f < F) bit = __ldg(X + r * ldx + f) >= 0.f;
    const unsigned word = __ballot_sync(0xffffffffu, bit);
    if (lane == 0) packed[item] = word;
}

// Vector path (X 16-byte aligned, ldx % 4 == 0): a warp owns U consecutive (row, 128-feature) items; each lane loads one
// float4 per item (512 contiguous bytes per warp-load, U of them in flight), turns it into a 4-bit nibble, and the 8
// lanes that share a 32-feature word OR their nibbles together with 3 xor-shuffles.  A streaming read: HBM bound.
// (The first version — one warp per (row, word), one 128-byte load per warp, ballot — reached 1.2 TB/s.)
template <int U>
__global__ void __launch_bounds__(256)
binarize_pack_vec_kernel(const float* __restrict__ X, long long ldx, uint32_t* __restrict__ packed,
                         long long wpr, long long n_items, int items_per_row, int F) {
    const int lane = threadIdx.x & 31;
    const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    for (long long it0 = warp * U; it0 < n_items; it0 += n_warps * U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long it = it0 + u;
            v[u] = make_float4(-1.f, -1.f, -1.f, -1.f);
            if (it < n_items) {
                const long long r = it / items_per_row;
                const int f = static_cast<int>(it - r * items_per_row) * 128 + lane * 4;
                const float* src = X + r * ldx + f;
                if (f + 4 <= F) {
                    v[u] = ldg_nc_f4(src);
                } else if (f < F) {                 // ragged tail of the row: 1..3 real features
                    v[u].x = __ldg(src);
                    if (f + 1 < F) v[u].y = __ldg(src + 1);
                    if (f + 2 < F) v[u].z = __ldg(src + 2);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long it = it0 + u;
            uint32_t nib = (v[u].x >= 0.f ? 1u : 0u) | (v[u].y >= 0.f ? 2u : 0u) | (v[u].z >= 0.f ? 4u : 0u) |
                           (v[u].w >= 0.f ? 8u : 0u);
            uint32_t w = nib << (4 * (lane & 7));
            w |= __shfl_xor_sync(0xffffffffu, w, 1);
            w |= __shfl_xor_sync(0xffffffffu, w, 2);
            w |= __shfl_xor_sync(0xffffffffu, w, 4);
            if (it < n_items && (lane & 7) == 0) {
                const long long r = it / items_per_row;
                const int c = static_cast<int>(it - r * items_per_row);
                packed[r * wpr + c * 4 + (lane >> 3)] = w;
            }
        }
    }
}

// Row-wise vector path for F <= 1024 (IPR = 128-feature items per row, a template parameter): a warp takes TWO whole
// rows per iteration, so the 2*IPR loads of a lane are issued together (5 KB in flight per warp at F = 602) and no item
// index has to be divided back into (row, item) — the item kernel above paid two 64-bit divisions per 512-byte item and
// reached 2.5 TB/s (0.37 of the HBM peak) on the Reddit-shaped table.
template <int IPR>
__global__ void __launch_bounds__(256)
binarize_pack_rows_kernel(const float* __restrict__ X, long long ldx, uint32_t* __restrict__ packed, long long wpr,
                          long long n_rows, int F) {
    const int lane = threadIdx.x & 31;
    const long long warp = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    // lane masks of the 128-feature items: a vector that starts below F is loaded whole (rows are 16-byte multiples, so it
    // stays inside the row) and the bits of columns >= F are masked off — no per-item branches, the 2*IPR loads of a
    // lane issue back to back (with a ragged-tail branch per item the compiler serialised load -> test -> load: ncu
    // showed one full memory latency per item, 0.30 ms = 0.29 of the HBM peak)
    uint32_t keep[IPR];
#pragma unroll
    for (int i = 0; i < IPR; ++i) {
        const int left = F - (i * 128 + lane * 4);
        keep[i] = left >= 4 ? 0xFu : (left <= 0 ? 0u : ((1u << left) - 1u));
    }
    for (long long r0 = warp * 2; r0 < n_rows; r0 += n_warps * 2) {
        float4 v[2][IPR];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const long long r = min(r0 + q, n_rows - 1);             // an odd last row is loaded twice, stored once
            const float* __restrict__ row = X + r * ldx + lane * 4;
#pragma unroll
            for (int i = 0; i < IPR; ++i)
                v[q][i] = keep[i] ? ldg_nc_f4(row + i * 128) : make_float4(-1.f, -1.f, -1.f, -1.f);
        }
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const long long r = r0 + q;
#pragma unroll
            for (int i = 0; i < IPR; ++i) {
                const uint32_t nib = ((v[q][i].x >= 0.f ? 1u : 0u) | (v[q][i].y >= 0.f ? 2u : 0u) |
                                      (v[q][i].z >= 0.f ? 4u : 0u) | (v[q][i].w >= 0.f ? 8u : 0u)) & keep[i];
                uint32_t w = nib << (4 * (lane & 7));
                w |= __shfl_xor_sync(0xffffffffu, w, 1);
                w |= __shfl_xor_sync(0xffffffffu, w, 2);
                w |= __shfl_xor_sync(0xffffffffu, w, 4);
                if (r < n_rows && (lane & 7) == 0) packed[r * wpr + i * 4 + (lane >> 3)] = w;
            }
        }
    }
}

template <int IPR>
static void launch_pack_rows(const float* X, long long ldx, uint32_t* packed, long long wpr, long long n_rows, int F,
                             int sm_count, cudaStream_t st) {
    long long blocks = ((n_rows + 1) / 2 * 32 + 255) / 256;
    const long long cap = static_cast<long long>(sm_count) * 8;
    if (blocks > cap) blocks = cap;
    binarize_pack_rows_kernel<IPR><<<static_cast<unsigned>(blocks), 256, 0, st>>>(X, ldx, packed, wpr, n_rows, F);
}

// ---- bit-sliced (carry-save) formulation -------------------------------------------------------------------
// One warp per destination row; LANE l OWNS PACKED WORD l of every neighbour row, so a neighbour costs one
// coalesced load of its whole packed row (wpr*4 contiguous bytes) instead of 32 lanes touching 32 different rows.
// Counts are kept BIT-SLICED: plane p of a lane holds bit p of the running count of each of its 32 features, so
// adding a neighbour's 32 features is a handful of bitwise ops for all 32 at once:
//   8 neighbours are folded into planes 0..2 with 7 carry-save adders (2 LOP3 each) and one ripple add of the
//   "eights" word into planes 3..7 (counts up to 255); every 248 neighbours, and at the end of the row, the planes
//   are scattered to the
//   feature-major integer accumulators with shuffles (lane f extracts bit f of word i's planes).
// ~10 bitwise ops per neighbour per warp instead of ~16 shuffle/LOP3 per neighbour for the 32x32 transposes of the
// first version, and the loads are coalesced (first version: 11.4 ms on the Reddit-shaped graph, issue + L1-wavefront
// bound; profiles/r01_kernels_before_policy.jsonl).
#define DGLLB_CSA(a, b, c, sum, carry)            \
    do {                                          \
        const uint32_t _a = (a), _b = (b), _c = (c); \
        sum = _a ^ _b ^ _c;                       \
        carry = (_a & _b) | (_a & _c) | (_b & _c); \
    } while (0)

template <int FIRST>
__device__ __forceinline__ void ripple_add(uint32_t (&pl)[8], uint32_t carry) {
#pragma unroll
    for (int p = FIRST; p < 8; ++p) {
        const uint32_t t = pl[p] & carry;
        pl[p] ^= carry;
        carry = t;
    }
}

// NW4 = 128-bit words per packed row (words_per_row = 4*NW4 <= 32)
// HEAVY = false: warp w owns row w (rows longer than `chunk` edges are skipped when chunk > 0).
// HEAVY = true : warp w owns item w = (row, k) of the plan: edges [k*chunk, (k+1)*chunk) of a long row; partial
//                results are combined with integer / exact-in-fp32 atomics into the pre-zeroed row (integer sums
//                commute, so the result stays bit-exact and deterministic).
template <int NW4, bool HEAVY>
__global__ void __launch_bounds__(256)
bin_spmm_kernel(const void* row_ptr, int rp64, const int* __restrict__ col,
                const uint32_t* __restrict__ packed, long long wpr, void* out, long long ldo,
                long long n_dst, int F, int out_mode, int chunk, const int2* __restrict__ items,
                long long n_items) {
    constexpr int NW = NW4 * 4;
    const int lane = threadIdx.x & 31;
    const long long wid = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    long long row, beg, end;
    if (HEAVY) {
        if (wid >= n_items) return;
        const int2 it = items[wid];
        row = it.x;
        const long long rb = bin_rp(row_ptr, rp64, row), re = bin_rp(row_ptr, rp64, row + 1);
        beg = rb + static_cast<long long>(it.y) * chunk;
        end = min(re, beg + chunk);
    } else {
        row = wid;
        if (row >= n_dst) return;
        beg = bin_rp(row_ptr, rp64, row);
        end = bin_rp(row_ptr, rp64, row + 1);
        if (chunk > 0 && end - beg > chunk) return;  // done by the heavy items
    }
    const bool word_on = lane < wpr;
    const uint32_t* __restrict__ pk = packed + lane;
    const unsigned uwpr = static_cast<unsigned>(wpr);

    uint32_t pl[8];
#pragma unroll
    for (int p = 0; p < 8; ++p) pl[p] = 0u;
    int acc[NW];
#pragma unroll
    for (int i = 0; i < NW; ++i) acc[i] = 0;
    int pending = 0;  // neighbours folded into the planes since the last scatter

    auto scatter_planes = [&]() {
        // number of planes that can be non-zero after `pending` neighbours
        const int kused = 32 - __clz(pending);
#pragma unroll
        for (int i = 0; i < NW; ++i) {
            if (i < wpr) {
                int c = 0;
#pragma unroll
                for (int p = 0; p < 8; ++p) {
                    if (p < kused) {
                        const uint32_t v = __shfl_sync(0xffffffffu, pl[p], i);
                        c |= static_cast<int>((v >> lane) & 1u) << p;
                    }
                }
                acc[i] += c;
            }
        }
#pragma unroll
        for (int p = 0; p < 8; ++p) pl[p] = 0u;
        pending = 0;
    };

    // Harley-Seal style block of 8 neighbours: planes 0..2 (ones, twos, fours) are updated with 7 carry-save adders
    // (2 LOP3 each), the resulting "eights" word ripples into planes 3..7 — ~3 bitwise ops per neighbour.
    for (long long e0 = beg; e0 < end; e0 += 32) {
        const int n = static_cast<int>(min(32ll, end - e0));
        const int my_c = lane < n ? __ldg(col + e0 + lane) : -1;
#pragma unroll 1
        for (int g8 = 0; g8 < n; g8 += 8) {
            uint32_t x[8];
            if (n - g8 >= 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const unsigned c = static_cast<unsigned>(__shfl_sync(0xffffffffu, my_c, g8 + u));
                    x[u] = word_on ? __ldg(pk + static_cast<unsigned long long>(c) * uwpr) : 0u;
                }
            } else {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int c = __shfl_sync(0xffffffffu, my_c, (g8 + u) & 31);
                    x[u] = (g8 + u < n && word_on)
                               ? __ldg(pk + static_cast<unsigned long long>(static_cast<unsigned>(c)) * uwpr) : 0u;
                }
            }
            uint32_t ta, tb, fa, fb, eights;
            DGLLB_CSA(pl[0], x[0], x[1], pl[0], ta);
            DGLLB_CSA(pl[0], x[2], x[3], pl[0], tb);
            DGLLB_CSA(pl[1], ta, tb, pl[1], fa);
            DGLLB_CSA(pl[0], x[4], x[5], pl[0], ta);
            DGLLB_CSA(pl[0], x[6], x[7], pl[0], tb);
            DGLLB_CSA(pl[1], ta, tb, pl[1], fb);
            DGLLB_CSA(pl[2], fa, fb, pl[2], eights);
            ripple_add<3>(pl, eights);
            pending += 8;  // (zero-padded slots add nothing; the bound stays valid)
            if (pending > 255 - 8) scatter_planes();
        }
    }
    if (pending) scatter_planes();

    const long long deg = end - beg;  // (HEAVY: edges of this chunk)
    const float fdeg = static_cast<float>(deg);
    const float inv = deg > 0 ? 1.f / fdeg : 0.f;
#pragma unroll
    for (int i = 0; i < NW; ++i) {
        const int f = i * 32 + lane;
        if (i < wpr && f < F) {
            if (out_mode == 0) {
                int* o = reinterpret_cast<int*>(out) + row * ldo + f;
                if (HEAVY) atomicAdd(o, acc[i]); else *o = acc[i];
            } else {
                const float sgn = 2.f * static_cast<float>(acc[i]) - fdeg;  // exact: |s| <= deg < 2^24 for our sizes
                float* o = reinterpret_cast<float*>(out) + row * ldo + f;
                if (HEAVY) atomicAdd(o, sgn);                 // integer-valued partial sums: exact, order independent
                else *o = out_mode == 1 ? sgn : sgn * inv;
            }
        }
    }
}

// heavy rows: zero before the atomics land / divide the +-1 sum by the row degree (mean mode)
__global__ void bin_heavy_zero_kernel(const int* __restrict__ heavy_rows, long long n_heavy, uint32_t* out,
                                      long long ldo, int F) {
    const long long r = blockIdx.x;
    if (r >= n_heavy) return;
    uint32_t* o = out + static_cast<long long>(heavy_rows[r]) * ldo;
    for (int f = threadIdx.x; f < F; f += blockDim.x) o[f] = 0u;
}
__global__ void bin_heavy_mean_kernel(const void* row_ptr, int rp64, const int* __restrict__ heavy_rows,
                                      long long n_heavy, float* out, long long ldo, int F) {
    const long long r = blockIdx.x;
    if (r >= n_heavy) return;
    const long long row = heavy_rows[r];
    const float inv = 1.f / static_cast<float>(bin_rp(row_ptr, rp64, row + 1) - bin_rp(row_ptr, rp64, row));
    float* o = out + row * ldo;
    for (int f = threadIdx.x; f < F; f += blockDim.x) o[f] *= inv;
}

}  // namespace dgllb

using namespace dgllb;

extern "C" int dgllb_binarize_pack(const float* X, int64_t ldx, uint32_t* packed,
                                   int64_t words_per_row, int64_t n_rows, int F, void* stream) {
    DGLLB_REQUIRE(n_rows >= 0 && F >= 0, "binarize_pack: negative size");
    if (n_rows == 0 || words_per_row == 0) return DGLLB_OK;
    DGLLB_REQUIRE(X && packed, "binarize_pack: null pointer");
    DGLLB_REQUIRE(words_per_row * 32 >= F && ldx >= F, "binarize_pack: words_per_row*32 < F or ldx < F");
    if (aligned16(X) && ldx % 4 == 0 && words_per_row % 4 == 0) {
        DevInfo di;
        int rc = get_devinfo(&di);
        if (rc != DGLLB_OK) return rc;
        constexpr int U = 4;
        const int ipr = static_cast<int>(words_per_row / 4);
        if (ipr <= 8) {
            cudaStream_t st = static_cast<cudaStream_t>(stream);
            switch (ipr) {
                case 1: launch_pack_rows<1>(X, ldx, packed, words_per_row, n_rows, F, di.sm_count, st); break;
                case 2: launch_pack_rows<2>(X, ldx, packed, words_per_row, n_rows, F, di.sm_count, st); break;
                case 3: launch_pack_rows<3>(X, ldx, packed, words_per_row, n_rows, F, di.sm_count, st); break;
                case 4: launch_pack_rows<4>(X, ldx, packed, words_per_row, n_rows, F, di.sm_count, st); break;
                case 5: launch_pack_rows<5>(X, ldx, packed, words_per_row, n_rows, F, di.sm_count, st); break;
                case 6: launch_pack_rows<6>(X, ldx, packed, words_per_row, n_rows, F, di.sm_count, st); break;
                case 7: launch_pack_rows<7>(X, ldx, packed, words_per_row, n_rows, F, di.sm_count, st); break;
                default: launch_pack_rows<8>(X, ldx, packed, words_per_row, n_rows, F, di.sm_count, st); break;
            }
            DGLLB_LAUNCH_CHECK();
            return DGLLB_OK;
        }
        const long long n_items = n_rows * ipr;
        long long blocks = ((n_items + U - 1) / U * 32 + 255) / 256;
        const long long cap = static_cast<long long>(di.sm_count) * 8 * 2;   // 8 resident CTAs per SM, two waves
        if (blocks > cap) blocks = cap;
        binarize_pack_vec_kernel<U><<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
            X, ldx, packed, words_per_row, n_items, ipr, F);
        DGLLB_LAUNCH_CHECK();
        return DGLLB_OK;
    }
    const long long items = n_rows * words_per_row;
    const long long blocks = (items * 32 + 255) / 256;
    DGLLB_REQUIRE(blocks < (1ll << 31), "binarize_pack: grid too large");
    binarize_pack_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        X, ldx, packed, words_per_row, n_rows, F);
    DGLLB_LAUNCH_CHECK();
    return DGLLB_OK;
}

extern "C" int dgllb_bin_spmm_csr(const void* row_ptr, int row_ptr_is64, const int32_t* col_idx,
                                  const uint32_t* packed, int64_t words_per_row, void* out,
                                  int64_t ldo, int64_t n_dst, int F, int out_mode,
                                  const dgllb_csr_plan* plan, void* stream) {
    DGLLB_REQUIRE(n_dst >= 0 && F >= 0, "bin_spmm: negative size");
    if (n_dst == 0 || F == 0) return DGLLB_OK;
    DGLLB_REQUIRE(row_ptr && col_idx && packed && out, "bin_spmm: null pointer");
    DGLLB_REQUIRE(out_mode >= 0 && out_mode <= 2, "bin_spmm: unknown out_mode %d", out_mode);
    DGLLB_REQUIRE(ldo >= F, "bin_spmm: ldo < F");
    DGLLB_REQUIRE(words_per_row % 4 == 0 && words_per_row * 32 >= F && aligned16(packed),
                  "bin_spmm: packed rows must be 16-byte aligned multiples of 4 words covering F "
                  "(words_per_row=%lld F=%d)", (long long)words_per_row, F);
    DGLLB_REQUIRE(!plan || plan->n_rows == n_dst, "bin_spmm: plan was built for another row count");
    const int nw4 = static_cast<int>(words_per_row / 4);
    if (nw4 > 8) {
        set_error("bin_spmm: F=%d needs %d x 128-bit words per row; this build supports <= 8 (F <= 1024)", F, nw4);
        return DGLLB_ERR_UNSUPPORTED;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool heavy = plan && plan->n_heavy_rows > 0;
    const int chunk = heavy ? plan->chunk_edges : 0;
    if (heavy) {
        bin_heavy_zero_kernel<<<static_cast<unsigned>(plan->n_heavy_rows), 128, 0, st>>>(
            plan->heavy_rows, plan->n_heavy_rows, static_cast<uint32_t*>(out), ldo, F);
        DGLLB_LAUNCH_CHECK();
    }
    int tb = opt_get(OPT_BIN_TB);  // default 64: small blocks retire evenly on ragged rows (4.54 -> 4.26 ms, Reddit-shaped)
    if (!(tb == 32 || tb == 64 || tb == 128 || tb == 256)) tb = 64;
    const long long blocks = (n_dst * 32 + tb - 1) / tb;
    const long long hblocks = heavy ? (plan->n_items * 32 + tb - 1) / tb : 0;
    DGLLB_REQUIRE(blocks < (1ll << 31) && hblocks < (1ll << 31), "bin_spmm: grid too large");
    const unsigned g = static_cast<unsigned>(blocks), hg = static_cast<unsigned>(hblocks);
#define DGLLB_BIN_CASE(NW)                                                                                   \
    case NW:                                                                                                 \
        if (heavy)                                                                                           \
            bin_spmm_kernel<NW, true><<<hg, tb, 0, st>>> (row_ptr, row_ptr_is64, col_idx, packed,            \
                                                          words_per_row, out, ldo, n_dst, F, out_mode,      \
                                                          chunk, plan->items, plan->n_items);               \
        bin_spmm_kernel<NW, false><<<g, tb, 0, st>>> (row_ptr, row_ptr_is64, col_idx, packed, words_per_row, \
                                                      out, ldo, n_dst, F, out_mode, chunk, nullptr, 0);      \
        break;
    switch (nw4) {
        DGLLB_BIN_CASE(1) DGLLB_BIN_CASE(2) DGLLB_BIN_CASE(3) DGLLB_BIN_CASE(4)
        DGLLB_BIN_CASE(5) DGLLB_BIN_CASE(6) DGLLB_BIN_CASE(7) DGLLB_BIN_CASE(8)
    }
#undef DGLLB_BIN_CASE
    g_launch_count.fetch_add(heavy ? 1 : 0);
    DGLLB_LAUNCH_CHECK();
    if (heavy && out_mode == 2) {
        bin_heavy_mean_kernel<<<static_cast<unsigned>(plan->n_heavy_rows), 128, 0, st>>>(
            row_ptr, row_ptr_is64, plan->heavy_rows, plan->n_heavy_rows, static_cast<float*>(out), ldo, F);
        DGLLB_LAUNCH_CHECK();
    }
    return DGLLB_OK;
}
