// binspmm.cu — bit-packed binarized neighbourhood aggregation (sm_100a).
//
// The reference only names this feature (README.md:11, "Quantization/Bineraztion"
// box in DGLL Architecture.drawio); semantics are fixed in SURVEY.md §8 a18:
//   X_b = (X >= 0) packed along the feature axis, 32 features per uint32 word;
//   cnt[i,f] = sum_{j in N(i)} X_b[j,f]           (int32, exact)
//   sum_pm1 = 2*cnt - deg_i ; mean_pm1 = sum_pm1 / deg_i  (fp32 epilogue)
//
// Popcount formulation: one warp per destination row, lanes = neighbours.
// Each lane fetches its neighbour's packed row with 128-bit loads; every
// 32-neighbour x 32-feature bit tile is transposed across the warp with five
// shuffle/LOP3 butterfly stages so lane f holds feature f's 32 neighbour bits,
// and __popc adds them to lane f's counter.  HBM traffic per edge is
// 4 B (index) + 4*words B (packed row): 76..80 B instead of 2,408 B at F=602.
// Algorithmic bytes: nnz*(4 + 4*ceil(F/32)) + n_dst*(4*F + r).
#include "common.cuh"

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

// 32x32 bit-matrix transpose across the lanes of a warp: before, lane i holds
// row i (bit b = element (i,b)); after, lane b holds column b (bit i = element (i,b)).
__device__ __forceinline__ uint32_t warp_bit_transpose(uint32_t x, int lane) {
#pragma unroll
    for (int s = 0; s < 5; ++s) {
        const int j = 16 >> s;
        const uint32_t m_lo = (s == 0) ? 0x0000FFFFu
                            : (s == 1) ? 0x00FF00FFu
                            : (s == 2) ? 0x0F0F0F0Fu
                            : (s == 3) ? 0x33333333u
                                       : 0x55555555u;
        const uint32_t y = __shfl_xor_sync(0xffffffffu, x, j);
        if (lane & j) x = (x & ~m_lo) | ((y >> j) & m_lo);
        else          x = (x & m_lo) | ((y << j) & ~m_lo);
    }
    return x;
}

// NW4 = 128-bit words per packed row (words_per_row = 4*NW4)
template <int NW4>
__global__ void __launch_bounds__(256)
bin_spmm_kernel(const void* row_ptr, int rp64, const int* __restrict__ col,
                const uint32_t* __restrict__ packed, long long wpr, void* out, long long ldo,
                long long n_dst, int F, int out_mode) {
    const int lane = threadIdx.x & 31;
    const long long row = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (row >= n_dst) return;
    const long long beg = bin_rp(row_ptr, rp64, row), end = bin_rp(row_ptr, rp64, row + 1);
    int cnt[NW4 * 4];
#pragma unroll
    for (int w = 0; w < NW4 * 4; ++w) cnt[w] = 0;

    for (long long e0 = beg; e0 < end; e0 += 32) {
        const bool on = e0 + lane < end;
        uint4 wv[NW4];
        if (on) {
            const uint4* src = reinterpret_cast<const uint4*>(packed + static_cast<long long>(__ldg(col + e0 + lane)) * wpr);
#pragma unroll
            for (int q = 0; q < NW4; ++q) wv[q] = __ldg(src + q);
        } else {
#pragma unroll
            for (int q = 0; q < NW4; ++q) wv[q] = make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int q = 0; q < NW4; ++q) {
            cnt[q * 4 + 0] += __popc(warp_bit_transpose(wv[q].x, lane));
            cnt[q * 4 + 1] += __popc(warp_bit_transpose(wv[q].y, lane));
            cnt[q * 4 + 2] += __popc(warp_bit_transpose(wv[q].z, lane));
            cnt[q * 4 + 3] += __popc(warp_bit_transpose(wv[q].w, lane));
        }
    }
    const long long deg = end - beg;
    const float fdeg = static_cast<float>(deg);
    const float inv = deg > 0 ? 1.f / fdeg : 0.f;
#pragma unroll
    for (int w = 0; w < NW4 * 4; ++w) {
        const int f = w * 32 + lane;
        if (f < F) {
            if (out_mode == 0) {
                reinterpret_cast<int*>(out)[row * ldo + f] = cnt[w];
            } else {
                const float s = 2.f * static_cast<float>(cnt[w]) - fdeg;  // exact: |s| <= deg < 2^24 for our sizes
                reinterpret_cast<float*>(out)[row * ldo + f] = out_mode == 1 ? s : s * inv;
            }
        }
    }
}

}  // namespace dgllb

using namespace dgllb;

extern "C" int dgllb_binarize_pack(const float* X, int64_t ldx, uint32_t* packed,
                                   int64_t words_per_row, int64_t n_rows, int F, void* stream) {
    DGLLB_REQUIRE(n_rows >= 0 && F >= 0, "binarize_pack: negative size");
    if (n_rows == 0 || words_per_row == 0) return DGLLB_OK;
    DGLLB_REQUIRE(X && packed, "binarize_pack: null pointer");
    DGLLB_REQUIRE(words_per_row * 32 >= F && ldx >= F, "binarize_pack: words_per_row*32 < F or ldx < F");
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
                                  int64_t ldo, int64_t n_dst, int F, int out_mode, void* stream) {
    DGLLB_REQUIRE(n_dst >= 0 && F >= 0, "bin_spmm: negative size");
    if (n_dst == 0 || F == 0) return DGLLB_OK;
    DGLLB_REQUIRE(row_ptr && col_idx && packed && out, "bin_spmm: null pointer");
    DGLLB_REQUIRE(out_mode >= 0 && out_mode <= 2, "bin_spmm: unknown out_mode %d", out_mode);
    DGLLB_REQUIRE(ldo >= F, "bin_spmm: ldo < F");
    DGLLB_REQUIRE(words_per_row % 4 == 0 && words_per_row * 32 >= F && aligned16(packed),
                  "bin_spmm: packed rows must be 16-byte aligned multiples of 4 words covering F "
                  "(words_per_row=%lld F=%d)", (long long)words_per_row, F);
    const int nw4 = static_cast<int>(words_per_row / 4);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long blocks = (n_dst * 32 + 255) / 256;
    DGLLB_REQUIRE(blocks < (1ll << 31), "bin_spmm: grid too large");
    const unsigned g = static_cast<unsigned>(blocks);
#define DGLLB_BIN_CASE(NW)                                                                              \
    case NW:                                                                                            \
        bin_spmm_kernel<NW><<<g, 256, 0, st>>>(row_ptr, row_ptr_is64, col_idx, packed, words_per_row,   \
                                               out, ldo, n_dst, F, out_mode);                           \
        break;
    switch (nw4) {
        DGLLB_BIN_CASE(1) DGLLB_BIN_CASE(2) DGLLB_BIN_CASE(3) DGLLB_BIN_CASE(4)
        DGLLB_BIN_CASE(5) DGLLB_BIN_CASE(6) DGLLB_BIN_CASE(7) DGLLB_BIN_CASE(8)
        default:
            set_error("bin_spmm: F=%d needs %d x 128-bit words per row; this build supports <= 8 (F <= 1024)",
                      F, nw4);
            return DGLLB_ERR_UNSUPPORTED;
    }
#undef DGLLB_BIN_CASE
    DGLLB_LAUNCH_CHECK();
    return DGLLB_OK;
}
