// spmm_bulk.cu — CSR aggregation with TMA-staged source rows (sm_100a).
//
// Same contract as spmm_stream_kernel (sum / mean over an explicit col_idx, row-aligned nnz-split, one owner
// warp per destination row, fp32 accumulation in CSR edge order, deterministic), different data movement:
//   * every WARP owns a ring of S shared-memory slots, one slot = one whole source feature row;
//   * lane 0 issues one cp.async.bulk (TMA engine, SASS UBLKCP) per edge: global row -> slot, completion
//     signalled on the slot's mbarrier (complete_tx::bytes) — no registers are tied up by loads in flight, so
//     an SM keeps warps*S*row_bytes (~190 KB at F=602) in flight instead of what the register file allows;
//   * all lanes wait on the slot's mbarrier, read the row with conflict-free LDS.128 (lane l takes 16-byte
//     vectors l, l+32, ...), accumulate, and the slot is re-armed for edge j+S.
// Per edge the warp issues ~1 bulk copy + NV LDS.128 + 4*NV FADD instead of NV * (SHFL + IMAD + LDG + ...) in
// the LDG kernels: the instruction stream is ~4x shorter, which matters because the LDG kernels are
// latency/issue bound at ~55% of the achievable gather bandwidth (profiles/r01_*).
// Algorithmic bytes per launch: nnz*(4 + [4] + F*b) + n_dst*(F*4 + r)   (SURVEY.md §8 d).
#include "common.cuh"
#include "internal.cuh"

namespace dgllb {

constexpr int kBulkWarps = 4;       // warps per CTA (several CTAs share an SM)
constexpr int kBulkMaxSlots = 8;    // ring depth per warp
constexpr int kBulkSmemBudget = 200 * 1024;  // per SM, leaves room for L1

__device__ __forceinline__ long long bulk_rp(const void* p, int is64, long long i) {
    return is64 ? reinterpret_cast<const long long*>(p)[i]
                : static_cast<long long>(reinterpret_cast<const int*>(p)[i]);
}

__device__ __forceinline__ float bulk_epi(float v, int epi) {
    if (epi & DGLLB_EPI_RELU) v = fmaxf(v, 0.f);
    if (epi & DGLLB_EPI_ELU) v = v > 0.f ? v : expm1f(v);
    return v;
}

template <bool BF16>
struct VecAcc;
template <>
struct VecAcc<false> {
    static constexpr int kA = 4;  // floats per 16-byte vector
    __device__ static __forceinline__ void add(float* acc, const uint4& r, float w, bool has_w) {
        const float x0 = __uint_as_float(r.x), x1 = __uint_as_float(r.y), x2 = __uint_as_float(r.z),
                    x3 = __uint_as_float(r.w);
        if (has_w) {
            acc[0] = fmaf(w, x0, acc[0]); acc[1] = fmaf(w, x1, acc[1]);
            acc[2] = fmaf(w, x2, acc[2]); acc[3] = fmaf(w, x3, acc[3]);
        } else {
            acc[0] += x0; acc[1] += x1; acc[2] += x2; acc[3] += x3;
        }
    }
};
template <>
struct VecAcc<true> {
    static constexpr int kA = 8;
    __device__ static __forceinline__ void add(float* acc, const uint4& r, float w, bool has_w) {
        const float x[8] = {bf16lo_to_f32(r.x), bf16hi_to_f32(r.x), bf16lo_to_f32(r.y), bf16hi_to_f32(r.y),
                            bf16lo_to_f32(r.z), bf16hi_to_f32(r.z), bf16lo_to_f32(r.w), bf16hi_to_f32(r.w)};
#pragma unroll
        for (int a = 0; a < 8; ++a) acc[a] = has_w ? fmaf(w, x[a], acc[a]) : acc[a] + x[a];
    }
};

// NV = 16-byte vectors per lane (row = up to 32*NV vectors)
template <bool BF16, int NV, bool HAS_VALS>
__global__ void __launch_bounds__(kBulkWarps * 32)
spmm_bulk_kernel(const SpmmParams p, const int T, const long long n_chunks, const int S, const int slot_bytes,
                 const int copy_bytes) {
    typedef VecAcc<BF16> V;
    constexpr int A = V::kA;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bars[kBulkWarps][kBulkMaxSlots];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long q = static_cast<long long>(blockIdx.x) * kBulkWarps + warp;
    if (q >= n_chunks) return;

    // rows owned by this warp: those whose first edge lies in [q*T, (q+1)*T)
    long long r0, r1;
    {
        const long long t = (q + (lane & 1)) * static_cast<long long>(T);
        long long lo = 0, hi = p.n_dst;
        while (lo < hi) {
            const long long mid = (lo + hi) >> 1;
            if (bulk_rp(p.row_ptr, p.rp64, mid) < t) lo = mid + 1; else hi = mid;
        }
        r0 = __shfl_sync(0xffffffffu, lo, 0);
        r1 = __shfl_sync(0xffffffffu, lo, 1);
        if (q == n_chunks - 1) r1 = p.n_dst;
    }
    if (r0 >= r1) return;
    const long long ebase = bulk_rp(p.row_ptr, p.rp64, r0);
    const int m = static_cast<int>(bulk_rp(p.row_ptr, p.rp64, r1) - ebase);

    unsigned char* ring = smem + static_cast<size_t>(warp) * S * slot_bytes;
    uint64_t* bar = bars[warp];
    if (lane == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();

    const char* __restrict__ Xc = reinterpret_cast<const char*>(p.X);
    const unsigned long long ld_bytes = static_cast<unsigned long long>(p.ldx) * (BF16 ? 2 : 4);
    const int* __restrict__ cb = p.col + ebase;
    const float* __restrict__ vb = HAS_VALS ? p.vals + ebase : nullptr;

    // column indices / values: 32 at a time, one chunk ahead
    int my_c = lane < m ? __ldg(cb + lane) : 0;
    int my_cn = 32 + lane < m ? __ldg(cb + 32 + lane) : 0;
    float my_w = 1.f, my_wn = 1.f;
    if (HAS_VALS) {
        my_w = lane < m ? __ldg(vb + lane) : 0.f;
        my_wn = 32 + lane < m ? __ldg(vb + 32 + lane) : 0.f;
    }
    long long row = r0, row_w0 = r0;
    int my_re = (row_w0 + lane < r1) ? static_cast<int>(bulk_rp(p.row_ptr, p.rp64, row_w0 + lane + 1) - ebase)
                                     : 0x7fffffff;
    int rstart = 0;
    int rend = __shfl_sync(0xffffffffu, my_re, 0);

    // prologue: arm the first S slots
    {
        const int n0 = min(S, m);
        for (int j = 0; j < n0; ++j) {
            const int c = __shfl_sync(0xffffffffu, my_c, j);
            if (lane == 0) {
                mbar_expect_tx(&bar[j], copy_bytes);
                bulk_g2s(ring + j * slot_bytes, Xc + static_cast<unsigned long long>(static_cast<unsigned>(c)) * ld_bytes,
                         copy_bytes, &bar[j]);
            }
        }
    }

    float acc[NV][A];
#pragma unroll
    for (int i = 0; i < NV; ++i)
#pragma unroll
        for (int a = 0; a < A; ++a) acc[i][a] = 0.f;

    const int n_vec = copy_bytes >> 4;  // 16-byte vectors per row
    const bool need_epi = p.row_scale || p.addend || p.bias || p.epi;

    auto flush = [&]() {
        const int deg = rend - rstart;
        const float scale = p.mean ? (deg > 0 ? 1.f / static_cast<float>(deg) : 0.f) : 1.f;
        float* __restrict__ orow = p.out + row * p.ldo;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int v = i * 32 + lane;
            const int c0 = v * A;
            if (v < n_vec && c0 < p.F) {
                if (p.out_vec && c0 + A <= p.F) {
#pragma unroll
                    for (int a = 0; a < A; a += 4) {
                        const float4 v4 = make_float4(acc[i][a] * scale, acc[i][a + 1] * scale,
                                                      acc[i][a + 2] * scale, acc[i][a + 3] * scale);
                        if (need_epi) *reinterpret_cast<float4*>(orow + c0 + a) = v4;
                        else stg_cs_f4(orow + c0 + a, v4);
                    }
                } else {
#pragma unroll
                    for (int a = 0; a < A; ++a)
                        if (c0 + a < p.F) orow[c0 + a] = acc[i][a] * scale;
                }
            }
#pragma unroll
            for (int a = 0; a < A; ++a) acc[i][a] = 0.f;
        }
        rstart = rend;
        ++row;
        if (row - row_w0 == 32) {
            row_w0 = row;
            my_re = (row_w0 + lane < r1)
                        ? static_cast<int>(bulk_rp(p.row_ptr, p.rp64, row_w0 + lane + 1) - ebase)
                        : 0x7fffffff;
        }
        rend = __shfl_sync(0xffffffffu, my_re, static_cast<int>(row - row_w0));
    };

    int slot = 0;
    uint32_t parity = 0;
    for (int j = 0; j < m; ++j) {
        while (j >= rend) flush();  // rows that ended before edge j (also empty rows)
        float w = 1.f;
        if (HAS_VALS) w = __shfl_sync(0xffffffffu, my_w, j & 31);
        mbar_wait(&bar[slot], parity);
        const uint4* src = reinterpret_cast<const uint4*>(ring + slot * slot_bytes);
        uint4 r[NV];
#pragma unroll
        for (int i = 0; i < NV; ++i)
            if (i * 32 + lane < n_vec) r[i] = src[i * 32 + lane];
#pragma unroll
        for (int i = 0; i < NV; ++i)
            if (i * 32 + lane < n_vec) V::add(acc[i], r[i], w, HAS_VALS);
        // index window bookkeeping: after consuming the last edge of a 32-chunk, slide
        const int jn = j + S;
        const int c = (jn >> 5) == (j >> 5) ? __shfl_sync(0xffffffffu, my_c, jn & 31)
                                            : __shfl_sync(0xffffffffu, my_cn, jn & 31);
        __syncwarp();  // every lane has consumed the slot (its values are in registers, accumulated)
        if (jn < m && lane == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&bar[slot], copy_bytes);
            bulk_g2s(ring + slot * slot_bytes, Xc + static_cast<unsigned long long>(static_cast<unsigned>(c)) * ld_bytes,
                     copy_bytes, &bar[slot]);
        }
        if ((j & 31) == 31) {
            my_c = my_cn;
            my_cn = j + 33 + lane < m ? __ldg(cb + j + 33 + lane) : 0;
            if (HAS_VALS) {
                my_w = my_wn;
                my_wn = j + 33 + lane < m ? __ldg(vb + j + 33 + lane) : 0.f;
            }
        }
        if (++slot == S) { slot = 0; parity ^= 1u; }
    }
    while (row < r1) flush();

    if (need_epi) {
        for (long long r = r0; r < r1; ++r) {
            float* __restrict__ orow = p.out + r * p.ldo;
            const float rs = p.row_scale ? __ldg(p.row_scale + r) : 1.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int c0 = (i * 32 + lane) * A;
#pragma unroll
                for (int a = 0; a < A; ++a) {
                    if (c0 + a < p.F) {
                        float v = orow[c0 + a] * rs;
                        if (p.addend) v += __ldg(p.addend + r * p.ld_add + c0 + a);
                        if (p.bias) v += __ldg(p.bias + c0 + a);
                        orow[c0 + a] = bulk_epi(v, p.epi);
                    }
                }
            }
        }
    }
}

template <bool BF16, int NV>
static int launch_bulk_nv(const SpmmParams& p, int T, long long n_chunks, int S, int slot_bytes, int copy_bytes,
                          cudaStream_t st) {
    const size_t smem = static_cast<size_t>(kBulkWarps) * S * slot_bytes;
    const long long blocks = (n_chunks + kBulkWarps - 1) / kBulkWarps;
    DGLLB_REQUIRE(blocks < (1ll << 31), "spmm: grid too large (%lld blocks)", blocks);
    if (p.vals) {
        DGLLB_CUDA_TRY(cudaFuncSetAttribute(spmm_bulk_kernel<BF16, NV, true>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        spmm_bulk_kernel<BF16, NV, true><<<static_cast<unsigned>(blocks), kBulkWarps * 32, smem, st>>>(
            p, T, n_chunks, S, slot_bytes, copy_bytes);
    } else {
        DGLLB_CUDA_TRY(cudaFuncSetAttribute(spmm_bulk_kernel<BF16, NV, false>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        spmm_bulk_kernel<BF16, NV, false><<<static_cast<unsigned>(blocks), kBulkWarps * 32, smem, st>>>(
            p, T, n_chunks, S, slot_bytes, copy_bytes);
    }
    DGLLB_LAUNCH_CHECK();
    return DGLLB_OK;
}

// returns DGLLB_ERR_UNSUPPORTED (without setting an error) when the shape does not fit this kernel
int spmm_bulk_try(const SpmmParams& p, int x_dtype, long long nnz, cudaStream_t st) {
    const bool bf16 = x_dtype == DGLLB_BF16;
    const int esz = bf16 ? 2 : 4;
    const long long row_bytes = static_cast<long long>(p.F) * esz;
    const int copy_bytes = static_cast<int>((row_bytes + 15) & ~15ll);
    if (!aligned16(p.X) || (p.ldx * esz) % 16 != 0 || copy_bytes > p.ldx * esz) return DGLLB_ERR_UNSUPPORTED;
    const int n_vec = copy_bytes / 16;
    const int nv = (n_vec + 31) / 32;
    if (nv > 8 || n_vec < 8) return DGLLB_ERR_UNSUPPORTED;
    DevInfo di;
    int rc = get_devinfo(&di);
    if (rc != DGLLB_OK) return rc;
    const int slot_bytes = (copy_bytes + 127) & ~127;
    // ring depth: as deep as the per-SM budget allows with >= 4 CTAs resident, capped at kBulkMaxSlots
    int S = kBulkSmemBudget / (4 * kBulkWarps * slot_bytes);
    if (S > kBulkMaxSlots) S = kBulkMaxSlots;
    if (S < 3) return DGLLB_ERR_UNSUPPORTED;
    // edges per warp: ~4 waves of (SMs x resident warps), between 32 and 1024
    const size_t smem_cta = static_cast<size_t>(kBulkWarps) * S * slot_bytes;
    int ctas_per_sm = static_cast<int>(kBulkSmemBudget / smem_cta);
    if (ctas_per_sm > 8) ctas_per_sm = 8;
    const long long target_warps = static_cast<long long>(di.sm_count) * ctas_per_sm * kBulkWarps * 4;
    long long T = (nnz + target_warps - 1) / target_warps;
    T = (T + 31) / 32 * 32;
    if (T < 32) T = 32;
    if (T > 1024) T = 1024;
    const long long n_chunks = nnz > 0 ? (nnz + T - 1) / T : 1;
#define DGLLB_BULK_CASE(NV)                                                                                   \
    case NV:                                                                                                  \
        return bf16 ? launch_bulk_nv<true, NV>(p, static_cast<int>(T), n_chunks, S, slot_bytes, copy_bytes, st) \
                    : launch_bulk_nv<false, NV>(p, static_cast<int>(T), n_chunks, S, slot_bytes, copy_bytes, st);
    switch (nv) {
        DGLLB_BULK_CASE(1) DGLLB_BULK_CASE(2) DGLLB_BULK_CASE(3) DGLLB_BULK_CASE(4)
        DGLLB_BULK_CASE(5) DGLLB_BULK_CASE(6) DGLLB_BULK_CASE(7) DGLLB_BULK_CASE(8)
    }
#undef DGLLB_BULK_CASE
    return DGLLB_ERR_UNSUPPORTED;
}

}  // namespace dgllb
