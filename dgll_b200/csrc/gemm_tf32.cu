// gemm_tf32.cu — the dense per-layer transform on tcgen05 tensor cores straight from the fp32 tensors in HBM.
//
// Replaces torch.mm(x, W) (dgll/nn/Convolution/gcnconv.py:30, gatconv.py:117, Evaluation/PPI/gcn_model.py:70) and its
// autograd products dX = G W^T, dW = X^T G when the caller asks for precision = 2 (TF32 operands, fp32 accumulation in
// TMEM; <= 2e-3 of max|ref|).  Round 1's tensor-core path (gemm_tcgen05.cu, precision = 1) first PACKED both operands
// to bf16 with a separate launch; ncu showed the layer shapes bound by that pass (160,000 x 602 x 256: 0.385 ms, of
// which 0.24 ms packing; profiles/r01_gemm_tcgen05.txt).  Here nothing is converted or copied:
//   * TMA (cp.async.bulk.tensor.2d, 128B swizzle) loads fp32 boxes of A and B as they lie in memory; tcgen05.mma
//     kind::tf32 reads the fp32 words (10-bit mantissa, low bits ignored) — M=128, N=128, K=8 per instruction;
//   * an operand whose reduction axis is contiguous is loaded K-major (box 32 k x 128 rows); an operand stored the
//     other way round ([K, rows]: the X^T and G^T of a backward pass, or a [K, N] weight) is loaded MN-major (four
//     boxes of 32 rows x 32 k, TMA swizzle 128B_ATOM_32B = the one MN-major layout tcgen05 accepts for 32-bit operands)
//     and the instruction descriptor's a_major / b_major bit tells the tensor core — no transposition pass either;
//   * CTAs of the same 128-row block of A are adjacent in launch order (blockIdx.x walks N first), so the second
//     128-column tile finds A in L2 and HBM moves A once;
//   * deterministic split-K (fixed-order reduction kernel) for the long-reduction / few-tile shapes (dW = X^T G).
// Pipeline per CTA (one 128 x 128 tile; 2 CTAs per SM): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane),
// warp 2 = TMEM allocator, warps 4-7 = epilogue (tcgen05.ld 32x32b.x32 -> (+C) + bias -> ReLU/ELU -> 128-bit stores).
// An operand TMA cannot address (base not 16-byte aligned, or row stride not a multiple of 16 bytes — e.g. a [256, 602]
// nn.Linear weight) is first copied into an aligned fp32 workspace by align_copy_kernel (small operands in practice).
// Roofline: HBM for the layer shapes of this path (K <= 1,204, N <= 256): bytes = 4*(M*K + N*K + M*N); tensor pipe
// (TF32 = half the bf16 rate) only for large square products, where precision = 1 remains the faster choice.
#include "common.cuh"
#include "internal.cuh"
#include <cuda.h>
#include <mutex>

namespace dgllb {

namespace tf32 {

constexpr int BM = 128, BN = 128, BK = 32, kStages = 3;      // BK fp32 = 128 bytes = one swizzle row
constexpr int kThreads = 256;
constexpr int kStageA = BM * BK * 4, kStageB = BN * BK * 4;  // 16 KB each
constexpr int kTmemCols = 128;

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_holder, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_holder)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]^T, tf32 x tf32 -> fp32
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Shared-memory matrix descriptor (descriptor version 1).
//   K-major  : layout type 2 (SWIZZLE_128B): rows of 128 bytes along K, 8-row groups 1024 B apart -> SBO = 1024, LBO unused
//   MN-major : 32-bit operands have ONE legal MN-major layout, type 1 (SWIZZLE_128B with a 32-byte swizzle base; what TMA
//              writes with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): atoms of 32 fp32 along MN (128 B) x 4 k-rows (512 B);
//              the next 4 k-rows follow at SBO = 512 B (consecutive rows of one TMA box), the next 32 MN elements at
//              LBO = 4096 B (the next TMA box).  One K=8 instruction spans two such atoms.
__device__ __forceinline__ uint64_t make_desc(const void* smem_ptr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_u32(smem_ptr) & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(lbo_bytes >> 4) << 16;
    d |= static_cast<uint64_t>(sbo_bytes >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(layout) << 61;
    return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float epi_fn(float v, int epi) {
    if (epi & DGLLB_EPI_RELU) v = fmaxf(v, 0.f);
    if (epi & DGLLB_EPI_ELU) v = v > 0.f ? v : expm1f(v);
    return v;
}

// A_MN / B_MN: the operand is stored with its MN axis contiguous ([K, rows]) and loaded MN-major.
template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kThreads)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 float* __restrict__ C, long long ldc, long long M, int N, int n_tiles_n, int num_kb_total,
                 int kb_per_split, long long split_slab, const float* __restrict__ bias, int epi, int accumulate) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* smem_a = smem;
    unsigned char* smem_b = smem + kStages * kStageA;
    __shared__ uint64_t full_bar[kStages], empty_bar[kStages], tmem_full_bar;
    __shared__ uint32_t tmem_base_holder;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // N tiles of one 128-row block of A are adjacent in launch order: the second finds A in L2
    const int tile_n = blockIdx.x % n_tiles_n;
    const long long tile_m = blockIdx.x / n_tiles_n;
    const int m0 = static_cast<int>(tile_m * BM), n0 = tile_n * BN;
    const int kb0 = blockIdx.z * kb_per_split;
    const int num_kb = min(kb_per_split, num_kb_total - kb0);
    C += static_cast<long long>(blockIdx.z) * split_slab;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_b)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(&tmem_base_holder, kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_holder;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1u);
                mbar_expect_tx(&full_bar[s], kStageA + kStageB);
                const int k0 = (kb0 + kb) * BK;
                unsigned char* sa = smem_a + s * kStageA;
                unsigned char* sb = smem_b + s * kStageB;
                if (A_MN) {
#pragma unroll
                    for (int b = 0; b < BM / 32; ++b) tma_load_2d(sa + b * 4096, &tmap_a, m0 + b * 32, k0, &full_bar[s]);
                } else {
                    tma_load_2d(sa, &tmap_a, k0, m0, &full_bar[s]);
                }
                if (B_MN) {
#pragma unroll
                    for (int b = 0; b < BN / 32; ++b) tma_load_2d(sb + b * 4096, &tmap_b, n0 + b * 32, k0, &full_bar[s]);
                } else {
                    tma_load_2d(sb, &tmap_b, k0, n0, &full_bar[s]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D = f32 (bit 4), A = B = TF32 (2 at [7,10) and [10,13)), a_major bit 15, b_major
            // bit 16 (1 = MN-major), N/8 at [17,23), M/16 at [24,29)
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (A_MN ? (1u << 15) : 0u) | (B_MN ? (1u << 16) : 0u) |
                                   (static_cast<uint32_t>(BN >> 3) << 17) | (static_cast<uint32_t>(BM >> 4) << 24);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint64_t da = A_MN ? make_desc(smem_a + s * kStageA, 4096, 512, 1) : make_desc(smem_a + s * kStageA, 16, 1024, 2);
                const uint64_t db = B_MN ? make_desc(smem_b + s * kStageB, 4096, 512, 1) : make_desc(smem_b + s * kStageB, 16, 1024, 2);
#pragma unroll
                for (int k = 0; k < BK / 8; ++k) {
                    // 8 tf32 along K: K-major = 32 bytes inside the swizzled row (+2 in 16-byte units);
                    //                 MN-major = the next 8 k-rows (+1024 bytes = +64)
                    tc_mma_tf32(tmem_base, da + static_cast<uint64_t>((A_MN ? 64 : 2) * k),
                                db + static_cast<uint64_t>((B_MN ? 64 : 2) * k), idesc, (kb | k) != 0 ? 1u : 0u);
                }
                tc_commit(&empty_bar[s]);
            }
            tc_commit(&tmem_full_bar);
        }
    } else if (warp >= 4) {
        const int q = warp - 4;
        mbar_wait(&tmem_full_bar, 0);
        tc_fence_after();
        // Every TMA load of this CTA has landed and been consumed (tmem_full_bar follows the last MMA), so the pipeline
        // stages are free: each epilogue warp stages its 32 x 32 chunk there to turn "one thread = one row" (TMEM lane
        // layout: 32 rows x 16 B per store instruction) into coalesced stores (4 rows x 128 B per instruction).
        constexpr int kLdT = 36;                                   // floats per staged row: 16-byte aligned, conflict-light
        float* tile = reinterpret_cast<float*>(smem) + q * (32 * kLdT);
        const long long row0 = static_cast<long long>(m0) + q * 32;
        const bool vec_ok = (ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            const int cbase = n0 + c * 32;
            if (cbase >= N) break;
            uint32_t r[32];
            tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c * 32), r);
            if (vec_ok && cbase + 32 <= N) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(tile + lane * kLdT + j) =
                        make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                    __uint_as_float(r[j + 3]));
                __syncwarp();
                const int c4 = (lane & 7) * 4;
                float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (bias) {
                    b4.x = __ldg(bias + cbase + c4); b4.y = __ldg(bias + cbase + c4 + 1);
                    b4.z = __ldg(bias + cbase + c4 + 2); b4.w = __ldg(bias + cbase + c4 + 3);
                }
#pragma unroll
                for (int it = 0; it < 8; ++it) {
                    const int rr = it * 4 + (lane >> 3);
                    const long long grow = row0 + rr;
                    if (grow < M) {
                        float4 v = *reinterpret_cast<const float4*>(tile + rr * kLdT + c4);
                        float* dst = C + grow * ldc + cbase + c4;
                        if (accumulate) {
                            const float4 o = *reinterpret_cast<const float4*>(dst);
                            v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
                        }
                        v.x = epi_fn(v.x + b4.x, epi); v.y = epi_fn(v.y + b4.y, epi);
                        v.z = epi_fn(v.z + b4.z, epi); v.w = epi_fn(v.w + b4.w, epi);
                        *reinterpret_cast<float4*>(dst) = v;
                    }
                }
                __syncwarp();
            } else {
                const long long row = row0 + lane;
                if (row < M) {
                    float* crow = C + row * ldc;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int col = cbase + j;
                        if (col < N) {
                            float v = __uint_as_float(r[j]);
                            if (accumulate) v += crow[col];
                            if (bias) v += __ldg(bias + col);
                            crow[col] = epi_fn(v, epi);
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------------------------ persistent variant --
// The kernel above runs ONE tile per CTA: set-up, the first TMA round trip, the MMAs and the epilogue of a tile are
// serial, and for short reductions (K = 100: four k-blocks) that chain IS the tile's time — 2,449,029 x 100 x 256 took
// 1.50 ms against 0.54 ms of HBM time (profiles/r02_gemm.jsonl), and every second CTA re-read its A block from L2.
// Here one CTA per SM walks tiles t = blockIdx.x, + gridDim.x, ...:
//   * BN = 256 when N > 128: the whole N = 256 output row block in one pass, A is read once;
//   * the TMA producer runs ahead ACROSS tiles (the smem ring does not drain at a tile boundary);
//   * two accumulators in TMEM (2 x BN columns): the MMA warp starts tile t+1 while the epilogue warps drain tile t
//     (tmem_full / tmem_empty barriers), so stores to C overlap the next tile's loads and MMAs;
//   * the epilogue stages through its own shared memory (the ring is busy) for coalesced 128-byte row segments.
// Used when there is at least one tile per SM and no split-K; same operands, descriptors and results as above.
//   * eight epilogue warps (two per scheduler; a warp reads the TMEM lane quarter warp % 4 and every second 32-column
//     chunk) turn their 32 x 32 chunk into a 128B-swizzled shared-memory box and ONE lane hands it to TMA
//     (cp.async.bulk.tensor store, rows / columns past M / N clipped by the tensor map).  The first version of this
//     kernel used four epilogue warps and per-lane LDS + STG.128 like the kernel above: one warp per scheduler with a
//     ~400-instruction dependent chain per chunk made the EPILOGUE the bound — 1.85 ms on the K = 100 shape, slower than
//     the one-tile kernel (ncu: issue active 17 %, tensor pipe 8 %, DRAM 23 %; profiles/r02_gemm_persistent.txt).
//     C that TMA cannot address (unaligned, or accumulate = read-modify-write) takes the staged LDS/STG path.
//
// X3 = true ("3xTF32", precision 3): fp32-grade results on the tensor cores.  Every fp32 operand word v is split in shared
// memory into hi = tf32(v) (round to nearest) and lo = v - hi (exact in fp32); the product is accumulated in TMEM as
// lo*hi + hi*lo + hi*hi — three kind::tf32 MMAs per k-step, the dropped lo*lo term and the tensor core's truncation of lo
// are ~2^-22 relative.  Four CONVERTER warps sit between the TMA producer and the MMA warp: they wait for a stage to land,
// rewrite it in place as hi and write lo next to it (an elementwise pass, so the swizzled K-major / MN-major layouts are
// preserved byte for byte), fence the generic-proxy writes for the async proxy and arrive on the stage's conv barrier,
// which is what the MMA warp waits for.  Split-K (items = tiles x splits, partial tiles to a workspace, fixed-order
// reduction) lives in this kernel for X3 so that every shape of a training step takes it.  Four epilogue warps.
constexpr int kPWarps = 12;
constexpr int kPThreads = kPWarps * 32;
constexpr int kPStageBytes = 8 * 4096;                               // one 32 x 32 fp32 box per epilogue warp (<= 8)
template <int BNP, bool X3> struct PersistCfg {
    static constexpr int stage_b = BNP * BK * 4;
    static constexpr int stage = (kStageA + stage_b) * (X3 ? 2 : 1);   // X3: [A hi | A lo | B hi | B lo]
    static constexpr int stages = (192 * 1024) / stage;                 // 4 x 48 KB, 6 x 32 KB, or 3 x 64 KB (X3, BN = 128)
    static constexpr int epi_warps = X3 ? 4 : 8;
    static constexpr size_t smem = static_cast<size_t>(stages) * stage + kPStageBytes + 1024;
};

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tmap, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                     reinterpret_cast<uint64_t>(tmap)),
                 "r"(c0), "r"(c1), "r"(smem_u32(smem_src))
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// v - trunc_tf32(v): what the tensor core drops when it reads the fp32 word v as tf32 (0 for Inf / NaN, which stay in hi)
__device__ __forceinline__ float tf32_lo(float v) {
    const float d = v - __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    return (fabsf(v) <= 3.4028234664e38f) ? d : 0.f;
}

template <bool A_MN, bool B_MN, int BNP, bool X3>
__global__ void __launch_bounds__(kPThreads, 1)
gemm_tf32_persistent_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                            const __grid_constant__ CUtensorMap tmap_c, int c_by_tma, float* __restrict__ C,
                            long long ldc, long long M, int N, int n_tiles_n, long long n_tiles, int num_kb_total,
                            int splits, int kb_per_split, long long slab_rows, const float* __restrict__ bias, int epi,
                            int accumulate) {
    typedef PersistCfg<BNP, X3> Cfg;
    constexpr int S = Cfg::stages;
    constexpr int kEpiWarps = Cfg::epi_warps;
    constexpr int kStage = Cfg::stage;
    constexpr int kOffB = X3 ? 2 * kStageA : kStageA;                 // B (hi) inside a stage
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    float* stage_c = reinterpret_cast<float*>(smem + S * kStage);                      // 1 KB aligned (swizzle atom)
    __shared__ uint64_t full_bar[S], conv_bar[S], empty_bar[S], tmem_full_bar[2], tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_holder;
    const long long n_items = n_tiles * splits;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_b)) : "memory");
        if (c_by_tma) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_c)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&conv_bar[s], 4);                             // one arrival per converter warp (X3 only)
            mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full_bar[a], 1);
            mbar_init(&tmem_empty_bar[a], kEpiWarps);               // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    constexpr int kAccCols = X3 ? 2 * BNP : BNP;                      // X3: two accumulators per stage, k-blocks alternate
    if (warp == 2) tmem_alloc(&tmem_base_holder, 2 * kAccCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_holder;

    if (warp == 0) {
        if (lane == 0) {
            long long it = 0;
            for (long long w = blockIdx.x; w < n_items; w += gridDim.x) {
                const long long t = w / splits;
                const int kb0 = static_cast<int>(w - t * splits) * kb_per_split;
                const int num_kb = min(kb_per_split, num_kb_total - kb0);
                const int n0 = static_cast<int>(t % n_tiles_n) * BNP;
                const int m0 = static_cast<int>(t / n_tiles_n) * BM;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = static_cast<int>(it % S);
                    const uint32_t ph = static_cast<uint32_t>(it / S) & 1u;
                    mbar_wait(&empty_bar[s], ph ^ 1u);
                    mbar_expect_tx(&full_bar[s], kStageA + Cfg::stage_b);
                    const int k0 = (kb0 + kb) * BK;
                    unsigned char* sa = smem + s * kStage;
                    unsigned char* sb = sa + kOffB;
                    if (A_MN) {
#pragma unroll
                        for (int b = 0; b < BM / 32; ++b) tma_load_2d(sa + b * 4096, &tmap_a, m0 + b * 32, k0, &full_bar[s]);
                    } else {
                        tma_load_2d(sa, &tmap_a, k0, m0, &full_bar[s]);
                    }
                    if (B_MN) {
#pragma unroll
                        for (int b = 0; b < BNP / 32; ++b) tma_load_2d(sb + b * 4096, &tmap_b, n0 + b * 32, k0, &full_bar[s]);
                    } else {
                        tma_load_2d(sb, &tmap_b, k0, n0, &full_bar[s]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (A_MN ? (1u << 15) : 0u) | (B_MN ? (1u << 16) : 0u) |
                                   (static_cast<uint32_t>(BNP >> 3) << 17) | (static_cast<uint32_t>(BM >> 4) << 24);
            long long it = 0;
            int lt = 0;
            for (long long w = blockIdx.x; w < n_items; w += gridDim.x, ++lt) {
                const int kb0 = static_cast<int>(w % splits) * kb_per_split;
                const int num_kb = min(kb_per_split, num_kb_total - kb0);
                const int as = lt & 1;
                const uint32_t aph = static_cast<uint32_t>(lt >> 1) & 1u;
                mbar_wait(&tmem_empty_bar[as], aph ^ 1u);           // the epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d0 = tmem_base + static_cast<uint32_t>(as * kAccCols);
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    // The tensor core's fp32 accumulation truncates: a chain of n accumulating MMAs comes out low by about
                    // n * 2^-24 of its magnitude (measured: 4.4e-6 of max|ref| at K = 602 with one accumulator = 76 k-steps).
                    // X3 alternates two accumulators per k-block — two chains of half the length, added with
                    // round-to-nearest in the epilogue — and the host caps the k-blocks of one item.
                    const uint32_t tmem_d = X3 ? tmem_d0 + static_cast<uint32_t>((kb & 1) * BNP) : tmem_d0;
                    const int kb_first = X3 ? (kb >> 1) : kb;        // 0 on the first k-block that touches this accumulator
                    const int s = static_cast<int>(it % S);
                    const uint32_t ph = static_cast<uint32_t>(it / S) & 1u;
                    mbar_wait(X3 ? &conv_bar[s] : &full_bar[s], ph);
                    tc_fence_after();
                    unsigned char* sa = smem + s * kStage;
                    unsigned char* sb = sa + kOffB;
                    const uint64_t da = A_MN ? make_desc(sa, 4096, 512, 1) : make_desc(sa, 16, 1024, 2);
                    const uint64_t db = B_MN ? make_desc(sb, 4096, 512, 1) : make_desc(sb, 16, 1024, 2);
                    if (X3) {
                        const uint64_t da_lo = A_MN ? make_desc(sa + kStageA, 4096, 512, 1) : make_desc(sa + kStageA, 16, 1024, 2);
                        const uint64_t db_lo = B_MN ? make_desc(sb + Cfg::stage_b, 4096, 512, 1)
                                                    : make_desc(sb + Cfg::stage_b, 16, 1024, 2);
#pragma unroll
                        for (int k = 0; k < BK / 8; ++k) {
                            const uint64_t oa = static_cast<uint64_t>((A_MN ? 64 : 2) * k), ob = static_cast<uint64_t>((B_MN ? 64 : 2) * k);
                            tc_mma_tf32(tmem_d, da_lo + oa, db + ob, idesc, (kb_first | k) != 0 ? 1u : 0u);   // small terms first
                            tc_mma_tf32(tmem_d, da + oa, db_lo + ob, idesc, 1u);
                            tc_mma_tf32(tmem_d, da + oa, db + ob, idesc, 1u);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < BK / 8; ++k)
                            tc_mma_tf32(tmem_d, da + static_cast<uint64_t>((A_MN ? 64 : 2) * k),
                                        db + static_cast<uint64_t>((B_MN ? 64 : 2) * k), idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    tc_commit(&empty_bar[s]);
                }
                tc_commit(&tmem_full_bar[as]);
            }
        }
    } else if (warp >= 4 && warp < 4 + kEpiWarps) {
        const int ew = warp - 4;
        const int q = ew & 3;                                       // TMEM lane quarter this warp may read (= warp % 4)
        const int half = ew >> 2;                                   // it takes the chunks c = half, half + kStep, ...
        constexpr int kStep = kEpiWarps / 4;
        float* tile = stage_c + ew * 1024;
        const int sw = lane & 7;                                    // swizzle term of this lane's staged row
        int lt = 0;
        float* const C0 = C;
        for (long long w = blockIdx.x; w < n_items; w += gridDim.x, ++lt) {
            const long long t = w / splits;
            const long long zrow = (w - t * splits) * slab_rows;    // split z writes rows [z*slab_rows, ...) of the workspace
            C = C0 + zrow * ldc;
            const int as = lt & 1;
            const uint32_t aph = static_cast<uint32_t>(lt >> 1) & 1u;
            const int n0 = static_cast<int>(t % n_tiles_n) * BNP;
            const long long row0 = (t / n_tiles_n) * BM + q * 32;
            mbar_wait(&tmem_full_bar[as], aph);
            tc_fence_after();
            const uint32_t tacc = tmem_base + static_cast<uint32_t>(as * kAccCols) + (static_cast<uint32_t>(q * 32) << 16);
            const int kb0_e = static_cast<int>(w - t * splits) * kb_per_split;
            const bool two_acc = X3 && min(kb_per_split, num_kb_total - kb0_e) > 1;   // the second accumulator was written
            constexpr int NC = BNP / 32;
            int n_chunks = (N - n0 + 31) / 32;
            if (n_chunks > NC) n_chunks = NC;
            bool released = false;
#pragma unroll 1
            for (int c = half; c < n_chunks; c += kStep) {
                const int cbase = n0 + c * 32;
                uint32_t r[32];
                tmem_ld32(tacc + static_cast<uint32_t>(c * 32), r);
                if (two_acc) {
                    uint32_t r2[32];
                    tmem_ld32(tacc + static_cast<uint32_t>(BNP + c * 32), r2);
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __uint_as_float(r2[j]));
                }
                if (c + kStep >= n_chunks) {                        // this warp's part of the accumulator is read out
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[as])) : "memory");
                    released = true;
                }
                if (c_by_tma) {
                    if (bias || epi) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            float v = __uint_as_float(r[j]);
                            if (bias && cbase + j < N) v += __ldg(bias + cbase + j);
                            r[j] = __float_as_uint(epi_fn(v, epi));
                        }
                    }
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous box left smem
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(tile + lane * 32 + ((j ^ sw) << 2)) =
                            make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                        __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0 && row0 < M) tma_store_2d(&tmap_c, tile, cbase, static_cast<int>(zrow + row0));
                } else {
                    const bool vec_ok = (ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && cbase + 32 <= N;
                    if (vec_ok) {
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            *reinterpret_cast<float4*>(tile + lane * 32 + ((j ^ sw) << 2)) =
                                make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                            __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
                        __syncwarp();
                        const int jj = lane & 7, c4 = jj * 4;
                        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (bias) {
                            b4.x = __ldg(bias + cbase + c4); b4.y = __ldg(bias + cbase + c4 + 1);
                            b4.z = __ldg(bias + cbase + c4 + 2); b4.w = __ldg(bias + cbase + c4 + 3);
                        }
#pragma unroll
                        for (int i8 = 0; i8 < 8; ++i8) {
                            const int rr = i8 * 4 + (lane >> 3);
                            const long long grow = row0 + rr;
                            if (grow < M) {
                                float4 v = *reinterpret_cast<const float4*>(tile + rr * 32 + ((jj ^ (rr & 7)) << 2));
                                float* dst = C + grow * ldc + cbase + c4;
                                if (accumulate) {
                                    const float4 o = *reinterpret_cast<const float4*>(dst);
                                    v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
                                }
                                v.x = epi_fn(v.x + b4.x, epi); v.y = epi_fn(v.y + b4.y, epi);
                                v.z = epi_fn(v.z + b4.z, epi); v.w = epi_fn(v.w + b4.w, epi);
                                *reinterpret_cast<float4*>(dst) = v;
                            }
                        }
                    } else {
                        const long long row = row0 + lane;
                        if (row < M) {
                            float* crow = C + row * ldc;
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const int col = cbase + j;
                                if (col < N) {
                                    float v = __uint_as_float(r[j]);
                                    if (accumulate) v += crow[col];
                                    if (bias) v += __ldg(bias + col);
                                    crow[col] = epi_fn(v, epi);
                                }
                            }
                        }
                    }
                }
            }
            if (!released) {                                        // no chunk of this tile was this warp's
                tc_fence_before();
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[as])) : "memory");
            }
        }
        if (c_by_tma && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before exit
    } else if (X3 && warp >= 8) {
        // converters: 128 threads rewrite a landed stage as (hi, lo); thread t takes the 16-byte vectors t, t + 128, ...
        const int ct = threadIdx.x - 8 * 32;
        long long it = 0;
        for (long long w = blockIdx.x; w < n_items; w += gridDim.x) {
            const int kb0 = static_cast<int>(w % splits) * kb_per_split;
            const int num_kb = min(kb_per_split, num_kb_total - kb0);
            for (int kb = 0; kb < num_kb; ++kb, ++it) {
                const int s = static_cast<int>(it % S);
                const uint32_t ph = static_cast<uint32_t>(it / S) & 1u;
                mbar_wait(&full_bar[s], ph);
                unsigned char* sa = smem + s * kStage;
#pragma unroll
                for (int part = 0; part < 2; ++part) {
                    float4* hi = reinterpret_cast<float4*>(part == 0 ? sa : sa + kOffB);
                    constexpr int nA = kStageA / 16, nB = Cfg::stage_b / 16;
                    const int nv = part == 0 ? nA : nB;
                    float4* lo = hi + nv;
#pragma unroll 4
                    for (int i = ct; i < nv; i += 128) {
                        // The tensor core reads an fp32 word as tf32 by IGNORING its low 13 mantissa bits, so the landed word
                        // itself serves as hi = trunc_tf32(v); only lo = rn_tf32(v - hi) has to be written (the stage is
                        // shared-memory-bandwidth bound: TMA 32 KB in, the MMAs read 96 KB, the converters read 32 KB and, with
                        // hi left in place, write 32 KB instead of 64 KB; ncu profiles/r02_gemm_x3.txt).  lo is rounded HERE
                        // because the tensor core would truncate it too.
                        const float4 v = hi[i];
                        float4 l;
                        uint32_t u;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(tf32_lo(v.x))); l.x = __uint_as_float(u);
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(tf32_lo(v.y))); l.y = __uint_as_float(u);
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(tf32_lo(v.z))); l.z = __uint_as_float(u);
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(tf32_lo(v.w))); l.w = __uint_as_float(u);
                        lo[i] = l;
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the MMA
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&conv_bar[s])) : "memory");
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 2 * kAccCols);
}

// dst[r, 0:cols] = src[r, 0:cols], dst row stride ldd (a multiple of 4 floats), pad columns zero
__global__ void __launch_bounds__(256)
align_copy_kernel(const float* __restrict__ src, long long lds, float* __restrict__ dst, long long ldd, long long rows,
                  int cols) {
    const long long total = rows * ldd;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / ldd;
        const int c = static_cast<int>(i - r * ldd);
        dst[i] = c < cols ? __ldg(src + r * lds + c) : 0.f;
    }
}

__global__ void __launch_bounds__(256)
splitk_reduce_kernel(const float* __restrict__ part, long long slab, int splits, float* __restrict__ C, long long ldc,
                     long long M, int N, int ldp, const float* __restrict__ bias, int epi, int accumulate) {
    const long long total = M * N;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / N;
        const int c = static_cast<int>(i - r * N);
        float v = accumulate ? C[r * ldc + c] : 0.f;
        const float* p = part + r * ldp + c;
        for (int z = 0; z < splits; ++z) v += p[z * slab];
        if (bias) v += __ldg(bias + c);
        C[r * ldc + c] = epi_fn(v, epi);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, []() {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// fp32 tensor stored [outer, inner] row-major with row stride ld (elements).  Box = box_inner x box_outer, 128B swizzle,
// out-of-bounds elements read as zero (K and M/N tails).
static int make_tmap(CUtensorMap* map, const float* base, long long inner, long long outer, long long ld, int box_inner,
                     int box_outer, CUtensorMapSwizzle swizzle) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("gemm: cuTensorMapEncodeTiled is not available from the driver");
        return DGLLB_ERR_CUDA;
    }
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 4};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_outer)};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("gemm: cuTensorMapEncodeTiled failed (%d) inner=%lld outer=%lld ld=%lld", static_cast<int>(r), inner,
                  outer, ld);
        return DGLLB_ERR_CUDA;
    }
    return DGLLB_OK;
}

static bool tma_ok(const float* p, long long ld) {
    return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld % 4) == 0;
}

template <bool A_MN, bool B_MN>
static cudaError_t set_smem_attr(size_t smem) {
    return cudaFuncSetAttribute(gemm_tf32_kernel<A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                static_cast<int>(smem));
}

template <bool A_MN, bool B_MN>
static void launch(dim3 grid, size_t smem, cudaStream_t st, const CUtensorMap& ta, const CUtensorMap& tb, float* C,
                   long long ldc, long long M, int N, int n_tiles_n, int num_kb, int kb_per, long long slab,
                   const float* bias, int epi, int accumulate) {
    gemm_tf32_kernel<A_MN, B_MN><<<grid, kThreads, smem, st>>>(ta, tb, C, ldc, M, N, n_tiles_n, num_kb, kb_per, slab,
                                                               bias, epi, accumulate);
}

struct PersistArgs {
    int grid;
    cudaStream_t st;
    CUtensorMap ta, tb, tc;
    int c_by_tma;
    float* C;
    long long ldc, M;
    int N, n_tiles_n;
    long long n_tiles;
    int num_kb, splits, kb_per;
    long long slab_rows;
    const float* bias;
    int epi, accumulate;
};

template <bool A_MN, bool B_MN, int BNP, bool X3>
static cudaError_t launch_persistent(const PersistArgs& a) {
    static std::once_flag once[64];
    static cudaError_t err[64];
    int dev_id = 0;
    cudaGetDevice(&dev_id);
    const int slot = dev_id & 63;
    std::call_once(once[slot], [&]() {
        err[slot] = cudaFuncSetAttribute(gemm_tf32_persistent_kernel<A_MN, B_MN, BNP, X3>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(PersistCfg<BNP, X3>::smem));
    });
    if (err[slot] != cudaSuccess) return err[slot];
    gemm_tf32_persistent_kernel<A_MN, B_MN, BNP, X3><<<a.grid, kPThreads, PersistCfg<BNP, X3>::smem, a.st>>>(
        a.ta, a.tb, a.tc, a.c_by_tma, a.C, a.ldc, a.M, a.N, a.n_tiles_n, a.n_tiles, a.num_kb, a.splits, a.kb_per,
        a.slab_rows, a.bias, a.epi, a.accumulate);
    return cudaSuccess;
}

template <int BNP, bool X3>
static cudaError_t launch_persistent_mn(bool a_mn, bool b_mn, const PersistArgs& a) {
    if (a_mn && b_mn) return launch_persistent<true, true, BNP, X3>(a);
    if (a_mn) return launch_persistent<true, false, BNP, X3>(a);
    if (b_mn) return launch_persistent<false, true, BNP, X3>(a);
    return launch_persistent<false, false, BNP, X3>(a);
}

}  // namespace tf32

static int gemm_tf32_impl(const float* A, long long lda, int transA, const float* B, long long ldb, int transB, float* C,
                          long long ldc, long long M, long long N, long long K, const float* bias, int epi, int accumulate,
                          bool x3, cudaStream_t st) {
    using namespace tf32;
    if (K == 0 || N >= (1ll << 31) || K >= (1ll << 30) || M >= (1ll << 31)) {
        return gemm_simt(A, lda, transA, B, ldb, transB, C, ldc, M, N, K, bias, epi, accumulate, st);
    }
    DevInfo di;
    { int rc_ = get_devinfo(&di); if (rc_ != DGLLB_OK) return rc_; }
    // cuTensorMapEncodeTiled is a DRIVER call: it needs the primary context bound to the calling thread, which a fresh
    // thread (the autograd engine's backward thread) only gets from a context-requiring runtime call — this query is
    // one, costs nothing and is legal during stream capture
    { cudaStreamCaptureStatus cs_; DGLLB_CUDA_TRY(cudaStreamIsCapturing(st, &cs_)); }
    // operands TMA cannot address are copied into an aligned workspace first (stored shape: rows x cols)
    const long long a_rows = transA ? K : M, a_cols = transA ? M : K;
    const long long b_rows = transB ? N : K, b_cols = transB ? K : N;
    const bool fix_a = !tma_ok(A, lda), fix_b = !tma_ok(B, ldb);
    const long long lda2 = (a_cols + 3) / 4 * 4, ldb2 = (b_cols + 3) / 4 * 4;
    const size_t bytes_a = fix_a ? ((static_cast<size_t>(a_rows) * lda2 * 4 + 255) & ~static_cast<size_t>(255)) : 0;
    const size_t bytes_b = fix_b ? ((static_cast<size_t>(b_rows) * ldb2 * 4 + 255) & ~static_cast<size_t>(255)) : 0;
    char* ws = nullptr;
    float* part = nullptr;
    int rc = DGLLB_OK;
    do {
        if (bytes_a + bytes_b) {
            cudaError_t e = cudaMallocAsync(&ws, bytes_a + bytes_b, st);
            if (e != cudaSuccess) { set_error("gemm: %s", cudaGetErrorString(e)); rc = DGLLB_ERR_CUDA; break; }
            const long long cap = static_cast<long long>(di.sm_count) * 8;
            if (fix_a) {
                float* d = reinterpret_cast<float*>(ws);
                long long nb = (a_rows * lda2 + 255) / 256;
                align_copy_kernel<<<static_cast<unsigned>(nb > cap ? cap : nb), 256, 0, st>>>(A, lda, d, lda2, a_rows,
                                                                                             static_cast<int>(a_cols));
                g_launch_count.fetch_add(1);
                A = d;
                lda = lda2;
            }
            if (fix_b) {
                float* d = reinterpret_cast<float*>(ws + bytes_a);
                long long nb = (b_rows * ldb2 + 255) / 256;
                align_copy_kernel<<<static_cast<unsigned>(nb > cap ? cap : nb), 256, 0, st>>>(B, ldb, d, ldb2, b_rows,
                                                                                             static_cast<int>(b_cols));
                g_launch_count.fetch_add(1);
                B = d;
                ldb = ldb2;
            }
        }
        // A operand (rows = M): stored [M, K] -> K-major box 32 (k) x 128 (rows); stored [K, M] -> MN-major boxes 32 x 32
        // B operand (rows = N): stored [N, K] (transB) -> K-major; stored [K, N] -> MN-major
        const bool a_mn = transA != 0, b_mn = transB == 0;
        CUtensorMap ta, tb;
        const CUtensorMapSwizzle sw_k = CU_TENSOR_MAP_SWIZZLE_128B, sw_mn = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
        if ((rc = a_mn ? make_tmap(&ta, A, M, K, lda, 32, BK, sw_mn) : make_tmap(&ta, A, K, M, lda, BK, BM, sw_k)) != DGLLB_OK) break;
        // persistent kernel (one CTA per SM, two accumulators): TF32 when there is at least one tile per SM; 3xTF32 always
        // (with its own split-K for the long-reduction / few-tile shapes)
        {
            // BN = 256 reads A once per row block, but pads N to a multiple of 256: keep it only while that costs < 10 %
            // more columns than 128-wide tiles would (N = 602: 768 vs 640 columns -> 128).  3xTF32 stages hold hi and lo
            // of both operands: BN = 128 (3 stages of 64 KB).
            const long long pad256 = (N + 255) / 256 * 256, pad128 = (N + 127) / 128 * 128;
            const int bnp = (!x3 && N > 128 && pad256 * 10 <= pad128 * 11) ? 256 : 128;
            const long long p_tiles_m = (M + BM - 1) / BM;
            const int p_tiles_n = static_cast<int>((N + bnp - 1) / bnp);
            const long long p_tiles = p_tiles_m * p_tiles_n;
            const int gk = opt_get(OPT_GEMM_KERNEL);   // 3 pins the one-tile-per-CTA kernel, 4 the persistent one
            if (x3 || ((p_tiles >= di.sm_count || gk == 4) && gk != 3 && p_tiles_m < (1ll << 24))) {
                if (p_tiles_m >= (1ll << 24)) { rc = DGLLB_ERR_UNSUPPORTED; set_error("gemm: too many row blocks"); break; }
                if ((rc = b_mn ? make_tmap(&tb, B, N, K, ldb, 32, BK, sw_mn) : make_tmap(&tb, B, K, N, ldb, BK, bnp, sw_k)) != DGLLB_OK) break;
                PersistArgs pa;
                pa.st = st; pa.ta = ta; pa.tb = tb; pa.tc = ta;
                pa.M = M; pa.N = static_cast<int>(N); pa.n_tiles_n = p_tiles_n; pa.n_tiles = p_tiles;
                pa.num_kb = static_cast<int>((K + BK - 1) / BK);
                pa.splits = 1; pa.kb_per = pa.num_kb; pa.slab_rows = 0;
                pa.C = C; pa.ldc = ldc; pa.bias = bias; pa.epi = epi; pa.accumulate = accumulate;
                if (x3 && p_tiles * 2 <= di.sm_count && pa.num_kb >= 16) {      // deterministic split-K, as the one-tile kernel
                    long long want = di.sm_count / p_tiles;
                    if (want > 32) want = 32;
                    if (want > pa.num_kb / 8) want = pa.num_kb / 8;
                    // the tensor core's fp32 accumulation truncates (see the MMA warp): one item accumulates at most 16 k-blocks
                    // (two chains of 32 k-steps), the partial tiles are summed with round-to-nearest by the reduction kernel
                    // (dW = X^T G over 233 K rows: 455 splits, 300 MB of partial tiles)
                    long long need = (pa.num_kb + 15) / 16;
                    const long long ws_cap = (1ll << 30) / (p_tiles * BM * bnp * 4);   // at most ~1 GB of partial tiles
                    if (need > ws_cap) need = ws_cap;
                    if (want < need) want = need;
                    if (want >= 2) {
                        pa.kb_per = static_cast<int>((pa.num_kb + want - 1) / want);
                        pa.splits = (pa.num_kb + pa.kb_per - 1) / pa.kb_per;
                    }
                }
                long long ldp = 0;
                if (pa.splits > 1) {
                    ldp = (N + 3) / 4 * 4;
                    pa.slab_rows = p_tiles_m * BM;                             // whole row blocks: a slab never spills into the next
                    cudaError_t e = cudaMallocAsync(&part, sizeof(float) * static_cast<size_t>(pa.slab_rows) * ldp * pa.splits, st);
                    if (e != cudaSuccess) { set_error("gemm: %s", cudaGetErrorString(e)); rc = DGLLB_ERR_CUDA; break; }
                    pa.C = part; pa.ldc = ldp; pa.bias = nullptr; pa.epi = 0; pa.accumulate = 0;
                }
                // C goes out by TMA store when the tensor map can address it and nothing has to be read back
                pa.c_by_tma = (tma_ok(pa.C, pa.ldc) && !pa.accumulate) ? 1 : 0;
                if (pa.c_by_tma) {
                    const long long c_rows = pa.splits > 1 ? pa.slab_rows * pa.splits : M;
                    if ((rc = make_tmap(&pa.tc, pa.C, N, c_rows, pa.ldc, 32, 32, sw_k)) != DGLLB_OK) break;
                }
                const long long items = p_tiles * pa.splits;
                pa.grid = static_cast<int>(items < di.sm_count ? items : di.sm_count);
                cudaError_t e = x3 ? launch_persistent_mn<128, true>(a_mn, b_mn, pa)
                              : bnp == 256 ? launch_persistent_mn<256, false>(a_mn, b_mn, pa)
                                           : launch_persistent_mn<128, false>(a_mn, b_mn, pa);
                if (e == cudaSuccess) e = cudaGetLastError();
                if (e != cudaSuccess) { set_error("gemm: persistent launch failed: %s", cudaGetErrorString(e)); rc = DGLLB_ERR_CUDA; break; }
                g_launch_count.fetch_add(1);
                if (pa.splits > 1) {
                    long long rb = (M * N + 255) / 256;
                    if (rb > static_cast<long long>(di.sm_count) * 8) rb = static_cast<long long>(di.sm_count) * 8;
                    splitk_reduce_kernel<<<static_cast<unsigned>(rb), 256, 0, st>>>(part, pa.slab_rows * ldp, pa.splits, C, ldc, M,
                                                                                    static_cast<int>(N), static_cast<int>(ldp),
                                                                                    bias, epi, accumulate);
                    g_launch_count.fetch_add(1);
                    e = cudaGetLastError();
                    if (e != cudaSuccess) { set_error("gemm: %s", cudaGetErrorString(e)); rc = DGLLB_ERR_CUDA; }
                }
                break;
            }
        }
        if ((rc = b_mn ? make_tmap(&tb, B, N, K, ldb, 32, BK, sw_mn) : make_tmap(&tb, B, K, N, ldb, BK, BN, sw_k)) != DGLLB_OK) break;
        const size_t smem = static_cast<size_t>(kStages) * (kStageA + kStageB) + 1024;
        static std::once_flag attr_once[64];
        static cudaError_t attr_err[64];
        int dev_id = 0;
        cudaGetDevice(&dev_id);
        const int slot = dev_id & 63;
        std::call_once(attr_once[slot], [&]() {
            cudaError_t e = set_smem_attr<false, false>(smem);
            if (e == cudaSuccess) e = set_smem_attr<false, true>(smem);
            if (e == cudaSuccess) e = set_smem_attr<true, false>(smem);
            if (e == cudaSuccess) e = set_smem_attr<true, true>(smem);
            attr_err[slot] = e;
        });
        if (attr_err[slot] != cudaSuccess) {
            set_error("gemm: %s", cudaGetErrorString(attr_err[slot]));
            rc = DGLLB_ERR_CUDA;
            break;
        }
        const long long tiles_m = (M + BM - 1) / BM;
        const int tiles_n = static_cast<int>((N + BN - 1) / BN);
        const long long tiles = tiles_m * tiles_n;
        if (tiles >= (1ll << 31)) { rc = DGLLB_ERR_UNSUPPORTED; set_error("gemm: too many tiles"); break; }
        dim3 grid(static_cast<unsigned>(tiles), 1, 1);
        const int num_kb = static_cast<int>((K + BK - 1) / BK);
        int splits = 1, kb_per = num_kb;
        if (tiles * 2 <= di.sm_count && num_kb >= 16) {
            long long want = di.sm_count / tiles;
            if (want > 32) want = 32;
            if (want > num_kb / 8) want = num_kb / 8;
            if (want >= 2) {
                kb_per = static_cast<int>((num_kb + want - 1) / want);
                splits = (num_kb + kb_per - 1) / kb_per;
            }
        }
        float* out = C;
        long long ldo = ldc, slab = 0;
        const float* kb_bias = bias;
        int kepi = epi, kacc = accumulate;
        if (splits > 1) {
            const int ldp = static_cast<int>((N + 3) / 4 * 4);
            slab = M * ldp;
            cudaError_t e = cudaMallocAsync(&part, sizeof(float) * static_cast<size_t>(slab) * splits, st);
            if (e != cudaSuccess) { set_error("gemm: %s", cudaGetErrorString(e)); rc = DGLLB_ERR_CUDA; break; }
            grid.z = static_cast<unsigned>(splits);
            out = part; ldo = ldp; kb_bias = nullptr; kepi = 0; kacc = 0;
        }
        const int Ni = static_cast<int>(N);
        if (a_mn && b_mn) launch<true, true>(grid, smem, st, ta, tb, out, ldo, M, Ni, tiles_n, num_kb, kb_per, slab, kb_bias, kepi, kacc);
        else if (a_mn) launch<true, false>(grid, smem, st, ta, tb, out, ldo, M, Ni, tiles_n, num_kb, kb_per, slab, kb_bias, kepi, kacc);
        else if (b_mn) launch<false, true>(grid, smem, st, ta, tb, out, ldo, M, Ni, tiles_n, num_kb, kb_per, slab, kb_bias, kepi, kacc);
        else launch<false, false>(grid, smem, st, ta, tb, out, ldo, M, Ni, tiles_n, num_kb, kb_per, slab, kb_bias, kepi, kacc);
        g_launch_count.fetch_add(1);
        if (splits > 1) {
            long long rb = (M * N + 255) / 256;
            if (rb > static_cast<long long>(di.sm_count) * 8) rb = static_cast<long long>(di.sm_count) * 8;
            splitk_reduce_kernel<<<static_cast<unsigned>(rb), 256, 0, st>>>(part, slab, splits, C, ldc, M, Ni,
                                                                            static_cast<int>(ldo), bias, epi, accumulate);
            g_launch_count.fetch_add(1);
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { set_error("gemm: launch failed: %s", cudaGetErrorString(e)); rc = DGLLB_ERR_CUDA; }
    } while (0);
    if (part) cudaFreeAsync(part, st);
    if (ws) cudaFreeAsync(ws, st);
    return rc;
}

int gemm_tf32(const float* A, long long lda, int transA, const float* B, long long ldb, int transB, float* C,
              long long ldc, long long M, long long N, long long K, const float* bias, int epi, int accumulate,
              cudaStream_t st) {
    return gemm_tf32_impl(A, lda, transA, B, ldb, transB, C, ldc, M, N, K, bias, epi, accumulate, false, st);
}

int gemm_tf32x3(const float* A, long long lda, int transA, const float* B, long long ldb, int transB, float* C,
                long long ldc, long long M, long long N, long long K, const float* bias, int epi, int accumulate,
                cudaStream_t st) {
    return gemm_tf32_impl(A, lda, transA, B, ldb, transB, C, ldc, M, N, K, bias, epi, accumulate, true, st);
}

}  // namespace dgllb
