// gemm_tcgen05.cu — the dense per-layer transform X·W on Blackwell's 5th-generation tensor cores (sm_100a).
//
// Replaces torch.mm(x, W) (dgll/nn/Convolution/gcnconv.py:30, gatconv.py:117, Evaluation/PPI/gcn_model.py:70) and the
// per-edge recomputed transform of gcn_fused_kernel.cu:46-54 when the caller asks for precision = 1 (bf16 operands,
// fp32 accumulation; parity bar 1e-2).  precision = 0 stays on the exact-fp32 SIMT GEMM.
//
// Pipeline (one CTA per 128 x 128 output tile, 2 CTAs per SM so one tile's epilogue overlaps the other's main loop):
//   pre-pass   pack_bf16_kernel: fp32 operand (either orientation) -> bf16, K-major, K padded to a multiple of 64
//   warp 0     TMA producer: cp.async.bulk.tensor.2d (UTMALDG) of a 128x64 A box and a 128x64 B box per k-block into a
//              3-stage ring of 128B-swizzled shared memory, completion on the stage's "full" mbarrier
//   warp 1     MMA issuer: one elected lane issues 4 x tcgen05.mma.cta_group::1.kind::f16 (M=128, N=128, K=16) per
//              k-block, accumulating in TMEM; tcgen05.commit frees the smem stage / signals the epilogue
//   warp 2     TMEM allocator (128 columns = 128 lanes x 128 fp32 accumulators)
//   warps 4-7  epilogue: tcgen05.ld 32x32b.x32 (each warp its own 32-lane quarter) -> registers -> (+C) + bias ->
//              ReLU/ELU -> fp32 global stores
// Roofline: tensor pipe for compute (2*M*N*K flop against the measured cuBLAS bf16 peak) — for the layer shapes of this
// path (K <= 1024, N <= 256) the GEMM is bound by reading the fp32 activations from HBM, not by the tensor pipe.
#include "common.cuh"
#include "internal.cuh"
#include <cuda.h>
#include <mutex>

namespace dgllb {

constexpr int kBM = 128, kBN = 128, kBK = 64, kStages = 3;
constexpr int kGemmThreads = 256;
constexpr int kStageBytesA = kBM * kBK * 2, kStageBytesB = kBN * kBK * 2;
constexpr int kTmemCols = 128;

// ------------------------------------------------------------------ pre-pass --
// dst[r, k] (bf16, row stride Kp) = TR ? src[k, r] : src[r, k];  zero for k >= K.  32x32 tiles through shared memory.
struct PackOperand {
    const float* src;
    long long ld, rows;
    __nv_bfloat16* dst;
    int transposed;      // 1: src is [K, rows] (read src[k, r]); 0: src is [rows, K]
    unsigned blocks_x;   // ceil(rows / 32)
};

// One block packs a 32 (rows) x 64 (k) tile of one operand.
//   direct     : thread = (row, 8 consecutive k): two 16-byte loads -> one 16-byte store of 8 bf16 (vector path when the
//                source rows are 16-byte aligned), 256 B read / 128 B written per row per tile
//   transposed : the tile goes through shared memory: reads coalesced along rows (128 B), writes 128 B along k
__device__ __forceinline__ void pack_tile(const PackOperand& o, unsigned bx, unsigned by, int K, int Kp,
                                          float (*tile)[33]) {
    const long long r0 = static_cast<long long>(bx) * 32;
    const int k0 = by * 64;
    if (o.transposed) {
        const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
        for (int j = ty; j < 64; j += 8) {
            const int k = k0 + j;
            const long long r = r0 + tx;
            tile[j][tx] = (k < K && r < o.rows) ? __ldg(o.src + static_cast<long long>(k) * o.ld + r) : 0.f;
        }
        __syncthreads();
        // 32 rows x 64 k: thread (row = tid / 8, k-octet = tid % 8) writes 8 bf16 = 16 bytes
        const int row = threadIdx.x >> 3, oct = threadIdx.x & 7;
        const long long r = r0 + row;
        const int k = k0 + oct * 8;
        if (r < o.rows && k < Kp) {   // Kp is a multiple of 64: the whole octet is inside the padded row
            __align__(16) __nv_bfloat16 v[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) v[a] = __float2bfloat16(tile[oct * 8 + a][row]);
            *reinterpret_cast<uint4*>(o.dst + r * Kp + k) = *reinterpret_cast<const uint4*>(v);
        }
    } else {
        const int row = threadIdx.x >> 3, oct = threadIdx.x & 7;
        const long long r = r0 + row;
        const int k = k0 + oct * 8;
        if (r >= o.rows || k >= Kp) return;
        const float* src = o.src + r * o.ld + k;
        float f[8];
        if (k + 8 <= K && (o.ld & 3) == 0 && (reinterpret_cast<uintptr_t>(o.src) & 15) == 0) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(src));
            const float4 b = __ldg(reinterpret_cast<const float4*>(src + 4));
            f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
        } else {
#pragma unroll
            for (int a = 0; a < 8; ++a) f[a] = (k + a < K) ? __ldg(src + a) : 0.f;
        }
        __align__(16) __nv_bfloat16 v[8];
#pragma unroll
        for (int a = 0; a < 8; ++a) v[a] = __float2bfloat16(f[a]);
        *reinterpret_cast<uint4*>(o.dst + r * Kp + k) = *reinterpret_cast<const uint4*>(v);
    }
}

// BOTH operands in one launch: blocks [0, a.blocks_x) of the x dimension pack A, the rest pack B (the small GEMMs of a
// training step are launch-latency bound: 4.5 us per pack launch against ~1 us of work).
// dst[r, k] (bf16, row stride Kp) = transposed ? src[k, r] : src[r, k];  zero for k >= K.
__global__ void __launch_bounds__(256)
pack_bf16_kernel(const PackOperand a, const PackOperand b, int K, int Kp) {
    __shared__ float tile[64][33];
    if (blockIdx.x < a.blocks_x) pack_tile(a, blockIdx.x, blockIdx.y, K, Kp, tile);
    else pack_tile(b, blockIdx.x - a.blocks_x, blockIdx.y, K, Kp, tile);
}

// split-K: out = (accumulate ? out : 0) + sum_z partial[z] (fixed order: deterministic) + bias, then the epilogue
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
        if (epi & DGLLB_EPI_RELU) v = fmaxf(v, 0.f);
        if (epi & DGLLB_EPI_ELU) v = v > 0.f ? v : expm1f(v);
        C[r * ldc + c] = v;
    }
}

// ------------------------------------------------------------------ PTX bits --
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
// D[tmem] (+)= A[smem desc] . B[smem desc]^T, bf16 x bf16 -> fp32
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// shared-memory matrix descriptor: K-major, SWIZZLE_128B, 128-byte rows, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_smem_desc(const void* smem_ptr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_u32(smem_ptr) & 0x3FFFF) >> 4);  // start address, 16-byte units, bits [0,14)
    d |= static_cast<uint64_t>(1) << 16;                               // leading byte offset (unused for SW128 K-major)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;                       // stride byte offset: 8 rows x 128 B
    d |= static_cast<uint64_t>(1) << 46;                               // descriptor version (Blackwell)
    d |= static_cast<uint64_t>(2) << 61;                               // layout type: SWIZZLE_128B
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

__device__ __forceinline__ float gemm_epi(float v, int epi) {
    if (epi & DGLLB_EPI_RELU) v = fmaxf(v, 0.f);
    if (epi & DGLLB_EPI_ELU) v = v > 0.f ? v : expm1f(v);
    return v;
}

// ------------------------------------------------------------------ kernel ---
__global__ void __launch_bounds__(kGemmThreads)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    float* __restrict__ C, long long ldc, long long M, int N, int num_kb_total, int kb_per_split,
                    long long split_slab, const float* __restrict__ bias, int epi, int accumulate) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // 1024-byte alignment is required by the 128B swizzle atom
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* smem_a = smem;
    unsigned char* smem_b = smem + kStages * kStageBytesA;
    __shared__ uint64_t full_bar[kStages], empty_bar[kStages], tmem_full_bar;
    __shared__ uint32_t tmem_base_holder;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * kBN;
    // split-K: this CTA reduces k-blocks [kb0, kb0 + num_kb) into its own slab of the partial buffer
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
                mbar_expect_tx(&full_bar[s], kStageBytesA + kStageBytesB);
                tma_load_2d(smem_a + s * kStageBytesA, &tmap_a, (kb0 + kb) * kBK, m0, &full_bar[s]);
                tma_load_2d(smem_b + s * kStageBytesB, &tmap_b, (kb0 + kb) * kBK, n0, &full_bar[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D=f32, A=B=bf16, both K-major, N/8 at [17,23), M/16 at [24,29)
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(kBN >> 3) << 17) |
                                   (static_cast<uint32_t>(kBM >> 4) << 24);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint64_t da = make_smem_desc(smem_a + s * kStageBytesA);
                const uint64_t db = make_smem_desc(smem_b + s * kStageBytesB);
#pragma unroll
                for (int k = 0; k < kBK / 16; ++k) {
                    // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in 16-byte units
                    tc_mma_bf16(tmem_base, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc,
                                (kb | k) != 0 ? 1u : 0u);
                }
                tc_commit(&empty_bar[s]);  // smem stage reusable once these MMAs have read it
            }
            tc_commit(&tmem_full_bar);     // accumulator complete
        }
    } else if (warp >= 4) {
        const int q = warp - 4;            // TMEM lane quarter == warp id % 4
        mbar_wait(&tmem_full_bar, 0);
        tc_fence_after();
        const long long row = static_cast<long long>(m0) + q * 32 + lane;
        const bool vec_ok = (ldc % 4 == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0) &&
                            (!bias || (reinterpret_cast<uintptr_t>(bias) & 15) == 0);
#pragma unroll 1
        for (int c = 0; c < kBN / 32; ++c) {
            uint32_t r[32];
            tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c * 32), r);
            if (row < M) {
                float* crow = C + row * ldc;
                const int cbase = n0 + c * 32;
                if (vec_ok && cbase + 32 <= N) {
                    // 8 x 128-bit stores: the thread owns 128 contiguous bytes of its output row
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 v = make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]),
                                               __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
                        if (accumulate) {
                            const float4 o = *reinterpret_cast<const float4*>(crow + cbase + j);
                            v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
                        }
                        if (bias) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + cbase + j));
                            v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
                        }
                        v.x = gemm_epi(v.x, epi); v.y = gemm_epi(v.y, epi);
                        v.z = gemm_epi(v.z, epi); v.w = gemm_epi(v.w, epi);
                        *reinterpret_cast<float4*>(crow + cbase + j) = v;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int col = cbase + j;
                        if (col < N) {
                            float v = __uint_as_float(r[j]);
                            if (accumulate) v += crow[col];
                            if (bias) v += __ldg(bias + col);
                            crow[col] = gemm_epi(v, epi);
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

// ------------------------------------------------------------------ host -----
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

// bf16 [rows, Kp] row-major, box = 64 (K) x 128 (rows), 128B swizzle, OOB rows read as zero
static int make_tmap(CUtensorMap* map, const void* base, long long rows, int Kp) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("gemm: cuTensorMapEncodeTiled is not available from the driver");
        return DGLLB_ERR_CUDA;
    }
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(Kp), static_cast<cuuint64_t>(rows)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(Kp) * 2};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(kBK), static_cast<cuuint32_t>(kBM)};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("gemm: cuTensorMapEncodeTiled failed (%d) rows=%lld Kp=%d", static_cast<int>(r), rows, Kp);
        return DGLLB_ERR_CUDA;
    }
    return DGLLB_OK;
}

int gemm_tcgen05(const float* A, long long lda, int transA, const float* B, long long ldb, int transB, float* C,
                 long long ldc, long long M, long long N, long long K, const float* bias, int epi, int accumulate,
                 cudaStream_t st) {
    if (K == 0 || N >= (1ll << 31) || K >= (1ll << 30)) {
        // degenerate / oversized: the SIMT path handles it exactly
        return gemm_simt(A, lda, transA, B, ldb, transB, C, ldc, M, N, K, bias, epi, accumulate, st);
    }
    DevInfo di;
    { int rc_ = get_devinfo(&di); if (rc_ != DGLLB_OK) return rc_; }  // also configures the workspace pool
    const int Kp = static_cast<int>((K + kBK - 1) / kBK * kBK);
    const size_t bytes_a = (static_cast<size_t>(M) * Kp * 2 + 255) & ~static_cast<size_t>(255);
    const size_t bytes_b = (static_cast<size_t>(N) * Kp * 2 + 255) & ~static_cast<size_t>(255);
    char* ws = nullptr;
    DGLLB_CUDA_TRY(cudaMallocAsync(&ws, bytes_a + bytes_b, st));
    __nv_bfloat16* Ab = reinterpret_cast<__nv_bfloat16*>(ws);
    __nv_bfloat16* Bb = reinterpret_cast<__nv_bfloat16*>(ws + bytes_a);
    int rc = DGLLB_OK;
    float* part = nullptr;
    do {
        PackOperand pa, pb;
        // A operand: rows = M, K-major.  stored [M,K] -> direct; stored [K,M] (transA) -> transpose
        pa.src = A; pa.ld = lda; pa.rows = M; pa.dst = Ab; pa.transposed = transA ? 1 : 0;
        pa.blocks_x = static_cast<unsigned>((M + 31) / 32);
        // B operand: rows = N, K-major (= B^T).  stored [K,N] -> transpose; stored [N,K] (transB) -> direct
        pb.src = B; pb.ld = ldb; pb.rows = N; pb.dst = Bb; pb.transposed = transB ? 0 : 1;
        pb.blocks_x = static_cast<unsigned>((N + 31) / 32);
        const dim3 gp(pa.blocks_x + pb.blocks_x, static_cast<unsigned>(Kp / 64));
        if (gp.y > 65535u) { rc = DGLLB_ERR_UNSUPPORTED; set_error("gemm: K too large"); break; }
        pack_bf16_kernel<<<gp, 256, 0, st>>>(pa, pb, static_cast<int>(K), Kp);
        g_launch_count.fetch_add(1);
        CUtensorMap ta, tb;
        if ((rc = make_tmap(&ta, Ab, M, Kp)) != DGLLB_OK) break;
        if ((rc = make_tmap(&tb, Bb, N, Kp)) != DGLLB_OK) break;
        const size_t smem = static_cast<size_t>(kStages) * (kStageBytesA + kStageBytesB) + 1024;
        // the opt-in shared-memory size is a per-device function attribute: set it once per device, not per call
        static std::once_flag attr_once[64];
        static cudaError_t attr_err[64];
        int dev_id = 0;
        cudaGetDevice(&dev_id);
        const int slot = dev_id & 63;
        std::call_once(attr_once[slot], [&]() {
            attr_err[slot] = cudaFuncSetAttribute(gemm_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                  static_cast<int>(smem));
        });
        if (attr_err[slot] != cudaSuccess) {
            set_error("gemm: %s", cudaGetErrorString(attr_err[slot]));
            rc = DGLLB_ERR_CUDA;
            break;
        }
        dim3 grid(static_cast<unsigned>((M + kBM - 1) / kBM), static_cast<unsigned>((N + kBN - 1) / kBN), 1);
        if (grid.y > 65535u) { rc = DGLLB_ERR_UNSUPPORTED; set_error("gemm: N too large"); break; }
        // split-K when the output has too few tiles to occupy the SMs and the reduction is long (dW = X^T G of a
        // training step: 2 x 5 tiles, K = 11,264 -> 85 us on 10 SMs).  Partials go to a workspace and are summed in a
        // fixed order by splitk_reduce_kernel: deterministic, unlike atomics.
        const int num_kb = Kp / kBK;
        const long long tiles = static_cast<long long>(grid.x) * grid.y;
        int splits = 1, kb_per = num_kb;
        if (tiles * 2 <= di.sm_count && num_kb >= 8) {
            long long want = di.sm_count / tiles;
            if (want > 32) want = 32;
            if (want > num_kb / 4) want = num_kb / 4;
            if (want >= 2) {
                kb_per = static_cast<int>((num_kb + want - 1) / want);
                splits = (num_kb + kb_per - 1) / kb_per;
            }
        }
        if (splits > 1) {
            const int ldp = static_cast<int>((N + 3) / 4 * 4);
            const long long slab = M * ldp;
            cudaError_t e = cudaMallocAsync(&part, sizeof(float) * static_cast<size_t>(slab) * splits, st);
            if (e != cudaSuccess) { set_error("gemm: %s", cudaGetErrorString(e)); rc = DGLLB_ERR_CUDA; break; }
            grid.z = static_cast<unsigned>(splits);
            gemm_tcgen05_kernel<<<grid, kGemmThreads, smem, st>>>(ta, tb, part, ldp, M, static_cast<int>(N), num_kb,
                                                                  kb_per, slab, nullptr, 0, 0);
            long long rb = (M * N + 255) / 256;
            if (rb > static_cast<long long>(di.sm_count) * 8) rb = static_cast<long long>(di.sm_count) * 8;
            splitk_reduce_kernel<<<static_cast<unsigned>(rb), 256, 0, st>>>(part, slab, splits, C, ldc, M,
                                                                            static_cast<int>(N), ldp, bias, epi,
                                                                            accumulate);
            g_launch_count.fetch_add(2);
        } else {
            gemm_tcgen05_kernel<<<grid, kGemmThreads, smem, st>>>(ta, tb, C, ldc, M, static_cast<int>(N), num_kb,
                                                                  num_kb, 0, bias, epi, accumulate);
            g_launch_count.fetch_add(1);
        }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { set_error("gemm: launch failed: %s", cudaGetErrorString(e)); rc = DGLLB_ERR_CUDA; }
    } while (0);
    if (part) cudaFreeAsync(part, st);
    cudaFreeAsync(ws, st);
    return rc;
}

}  // namespace dgllb
