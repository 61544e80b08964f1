// membw.cu — ceilings for the aggregation kernels on this GPU (diagnostic tool, not part of the library):
//   1. streaming read of a 4 GiB array (HBM read bandwidth, 128-bit loads)
//   2. repeated read of a 32 MiB array (L2 -> SM fabric bandwidth)
//   3. random gather of 512-byte row pieces from a 563 MB table (the SpMM access pattern without any arithmetic
//      dependence: every lane keeps DEPTH independent 16-byte loads in flight)
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o membw membw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ float4 ldg_nc(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

template <int U>
__global__ void __launch_bounds__(256) stream_read(const float4* __restrict__ a, size_t n4, float* sink) {
    size_t i = (size_t)blockIdx.x * blockDim.x * U + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x * U;
    float s = 0.f;
    for (; i + (size_t)(U - 1) * blockDim.x < n4; i += stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = ldg_nc(a + i + (size_t)u * blockDim.x);
#pragma unroll
        for (int u = 0; u < U; ++u) s += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    if (s == 123.456f) *sink = s;
}

// each warp: for it in [0, iters): DEPTH random rows; lane reads 16 B at row*stride + piece*512 + lane*16
template <int DEPTH>
__global__ void __launch_bounds__(256) gather_read(const char* __restrict__ table, uint32_t n_rows, uint32_t stride,
                                                   int pieces, int iters, float* sink) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t state = warp * 2654435761u + 12345u;
    const int piece = warp % pieces;
    float s = 0.f;
    for (int it = 0; it < iters; ++it) {
        float4 v[DEPTH];
#pragma unroll
        for (int u = 0; u < DEPTH; ++u) {
            state = state * 1664525u + 1013904223u;
            const uint32_t row = (uint32_t)(((uint64_t)(state >> 4) * n_rows) >> 28);
            v[u] = ldg_nc(reinterpret_cast<const float4*>(table + (size_t)row * stride + piece * 512 + lane * 16));
        }
#pragma unroll
        for (int u = 0; u < DEPTH; ++u) s += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    if (s == 123.456f) *sink = s;
}

// whole-row variant (what spmm_rows_kernel<float,5,4> does): each warp keeps DEPTH random ROWS in flight, 5 x 512 B each
// (lane reads 16 B at row*stride + s*512 + lane*16, s = 0..4; the 5th piece is the ragged tail of a 2,408-byte row).
// skew > 0 draws a share of the rows from a small hot set, which reproduces the L2 hit rate of a sampled mini-batch.
template <int DEPTH>
__global__ void __launch_bounds__(64) gather_rows(const char* __restrict__ table, uint32_t n_rows, uint32_t stride,
                                                  int iters, uint32_t hot_rows, uint32_t hot_per_1024, float* sink) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t state = warp * 2654435761u + 12345u;
    float s = 0.f;
    const bool last_on = lane * 16 + 2048 < 2408;
    for (int it = 0; it < iters; ++it) {
        float4 v[DEPTH][5];
#pragma unroll
        for (int u = 0; u < DEPTH; ++u) {
            state = state * 1664525u + 1013904223u;
            uint32_t row = (uint32_t)(((uint64_t)(state >> 4) * n_rows) >> 28);
            if (((state >> 3) & 1023u) < hot_per_1024) row = row % hot_rows;
            const char* p = table + (size_t)row * stride + lane * 16;
#pragma unroll
            for (int q = 0; q < 5; ++q)
                if (q < 4 || last_on) v[u][q] = ldg_nc(reinterpret_cast<const float4*>(p + q * 512));
        }
#pragma unroll
        for (int u = 0; u < DEPTH; ++u)
#pragma unroll
            for (int q = 0; q < 5; ++q)
                if (q < 4 || last_on) s += v[u][q].x + v[u][q].y + v[u][q].z + v[u][q].w;
    }
    if (s == 123.456f) *sink = s;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b) { float ms; cudaEventElapsedTime(&ms, a, b); return ms; }

int main() {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float* sink; cudaMalloc(&sink, 4);
    const size_t big = (size_t)4 << 30;
    float4* a; cudaMalloc(&a, big); cudaMemset(a, 0, big);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        stream_read<8><<<sms * 8, 256>>>(a, big / 16, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        if (rep) printf("hbm_stream_read_4GiB      %8.1f GB/s\n", big / (time_ms(e0, e1) * 1e-3) / 1e9);
    }
    const size_t small = (size_t)32 << 20;
    for (int rep = 0; rep < 2; ++rep) {
        stream_read<8><<<sms * 8, 256>>>(a, small / 16, sink);  // warm L2
        cudaEventRecord(e0);
        for (int i = 0; i < 50; ++i) stream_read<8><<<sms * 8, 256>>>(a, small / 16, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        if (rep) printf("l2_resident_read_32MiB    %8.1f GB/s\n", 50.0 * small / (time_ms(e0, e1) * 1e-3) / 1e9);
    }
    const uint32_t n_rows = 232965, stride = 2416;
    const int iters = 64;
    for (int blocks_per_sm = 2; blocks_per_sm <= 8; blocks_per_sm *= 2) {
        const int blocks = sms * blocks_per_sm * 4;
        const double bytes = (double)blocks * 8 * (iters * 8) * 512.0;
#define RUN(D)                                                                                              \
        for (int rep = 0; rep < 2; ++rep) {                                                                 \
            cudaEventRecord(e0);                                                                            \
            gather_read<D><<<blocks, 256>>>((const char*)a, n_rows, stride, 5, iters / D * 8, sink);         \
            cudaEventRecord(e1); cudaEventSynchronize(e1);                                                  \
            if (rep) printf("gather_512B depth=%-2d blocks=%-5d %8.1f GB/s\n", D, blocks,                   \
                            bytes / (time_ms(e0, e1) * 1e-3) / 1e9);             \
        }
        RUN(4) RUN(8) RUN(16)
    }
    // whole 2,408-byte rows, DEPTH rows in flight per warp, 64-thread blocks (the launch shape of spmm_rows_kernel)
    for (int hot = 0; hot <= 400; hot += 400) {
        for (int warps_per_sm = 16; warps_per_sm <= 32; warps_per_sm *= 2) {
            const int blocks = sms * warps_per_sm / 2 * 4;   // 4 waves
            const int it4 = 16;
            const double bytes = (double)blocks * 2 * it4 * 4 * 2408.0;
#define RUNR(D)                                                                                              \
            for (int rep = 0; rep < 2; ++rep) {                                                                 \
                cudaEventRecord(e0);                                                                            \
                gather_rows<D><<<blocks, 64>>>((const char*)a, n_rows, stride, it4 * 4 / D, 20000, hot, sink);   \
                cudaEventRecord(e1); cudaEventSynchronize(e1);                                                  \
                if (rep) printf("gather_rows_2408B depth=%d warps/SM=%-2d hot=%d/1024  %8.1f GB/s\n", D, warps_per_sm, hot, \
                                bytes / (time_ms(e0, e1) * 1e-3) / 1e9);                                       \
            }
            RUNR(2) RUNR(4)
        }
    }
    // uniform random gather over a table that fits L2 (60 MB): the fabric ceiling for this access shape
    for (int rep = 0; rep < 2; ++rep) {
        const int blocks = sms * 16;
        cudaEventRecord(e0);
        gather_read<8><<<blocks, 256>>>((const char*)a, 24000, stride, 5, 64, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        if (rep) printf("gather_512B L2-resident   %8.1f GB/s\n", (double)blocks * 8 * 64 * 8 * 512.0 / (time_ms(e0, e1) * 1e-3) / 1e9);
    }
    cudaError_t err = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(err));
    return 0;
}
