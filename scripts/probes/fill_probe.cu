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
    return 0;
}
