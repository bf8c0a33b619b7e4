// Store-pattern probe: which write pattern reaches the memset bandwidth on B200?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_gridstride(float4* p, size_t n, int cs) {
    const float4 z = make_float4(0, 0, 0, 0);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (cs) __stcs(p + i, z); else p[i] = z;
    }
}
// each warp writes `per` consecutive 512-byte rows (a tile), then jumps n_warps tiles ahead
__global__ void k_warptile(float4* p, size_t n, int per, int cs) {
    const float4 z = make_float4(0, 0, 0, 0);
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    const size_t tile_units = (size_t)per * 32, tiles = n / tile_units;
    for (size_t t = warp; t < tiles; t += nw) {
        float4* d = p + t * tile_units;
        #pragma unroll 4
        for (int f = lane; f < (int)tile_units; f += 32) { if (cs) __stcs(d + f, z); else d[f] = z; }
    }
}
// one CTA writes a contiguous chunk (per*8 rows of 512 B), CTAs stride
__global__ void k_ctatile(float4* p, size_t n, int per, int cs) {
    const float4 z = make_float4(0, 0, 0, 0);
    const size_t tile_units = (size_t)per * blockDim.x, tiles = n / tile_units;
    for (size_t t = blockIdx.x; t < tiles; t += gridDim.x) {
        float4* d = p + t * tile_units;
        #pragma unroll 4
        for (int f = threadIdx.x; f < (int)tile_units; f += blockDim.x) { if (cs) __stcs(d + f, z); else d[f] = z; }
    }
}
// TMA bulk stores: a CTA keeps one zeroed tile in shared memory and one thread per warp issues
// cp.async.bulk.global.shared::cta copies of `tile` bytes (no LSU store instructions at all)
__global__ void k_bulk(char* p, size_t bytes, int tile, int persistent) {
    extern __shared__ __align__(128) unsigned char sm[];
    for (int i = threadIdx.x * 16; i < tile; i += blockDim.x * 16) *reinterpret_cast<float4*>(sm + i) = make_float4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const size_t tiles = bytes / tile;
    const int lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((size_t)gridDim.x * blockDim.x) >> 5;
    if (lane == 0) {
        const unsigned src = (unsigned)__cvta_generic_to_shared(sm);
        for (size_t t = warp; t < tiles; t += nw) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p + t * tile), "r"(src), "r"(tile) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (!persistent) break;
            asm volatile("cp.async.bulk.wait_group.read 8;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}
template <typename F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f(); cudaDeviceSynchronize();
    cudaEventRecord(a); for (int i = 0; i < 5; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / 5;
}
int main() {
    const size_t bytes = 10ull << 30, n = bytes / 16;
    float4* p; cudaMalloc(&p, bytes);
    auto gbs = [&](float ms) { return bytes / 1e9 / (ms / 1e3); };
    printf("cudaMemset                 %6.0f GB/s\n", gbs(timeit([&] { cudaMemsetAsync(p, 0, bytes); })));
    for (int cs = 0; cs < 2; ++cs) {
        for (int blocks : {148 * 8, 148 * 16, 148 * 64, 1 << 20})
            printf("gridstride cs=%d blocks=%7d %6.0f GB/s\n", cs, blocks, gbs(timeit([&] { k_gridstride<<<blocks, 256>>>(p, n, cs); })));
        for (int per : {1, 4, 11, 44})
            printf("warptile  cs=%d per=%2d        %6.0f GB/s\n", cs, per, gbs(timeit([&] { k_warptile<<<148 * 8, 256>>>(p, n, per, cs); })));
        for (int per : {4, 11, 44})
            printf("ctatile   cs=%d per=%2d        %6.0f GB/s\n", cs, per, gbs(timeit([&] { k_ctatile<<<148 * 8, 256>>>(p, n, per, cs); })));
    }
    for (int tile : {2048, 4096, 5632, 8192, 16384, 32768}) {
        cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, tile);
        for (int ctas : {148 * 2, 148 * 4, 148 * 8}) {
            printf("bulk persistent tile=%5d ctas=%5d   %6.0f GB/s\n", tile, ctas,
                   gbs(timeit([&] { k_bulk<<<ctas, 256, tile>>>((char*)p, bytes, tile, 1); })));
        }
        const size_t tiles = bytes / tile;
        const unsigned blocks = (unsigned)((tiles + 7) / 8);
        printf("bulk one-shot   tile=%5d blocks=%7u %6.0f GB/s\n", tile, blocks,
               gbs(timeit([&] { k_bulk<<<blocks, 256, tile>>>((char*)p, bytes, tile, 0); })));
    }
    { cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e)); }
    return 0;
}
