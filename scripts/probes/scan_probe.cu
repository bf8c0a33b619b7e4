// Access-pattern probe for the decoder's two kinds of reads on B200:
//   scan : 16 bytes (objectness + 3 anchor logits) out of every 352-byte row
//   rows : whole 352-byte rows of a random ~18% of the rows, into shared memory
// Which issue mechanism / how much in flight does it take to reach the DRAM limit of
// each pattern, and how many bytes does each really pull (use ncu for dram__bytes)?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scan_probe scan_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int D = 88;   // floats per row

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// load flavours for the 16-byte scan load
template <int V> __device__ __forceinline__ float4 ld16(const float* p)
{
    float4 v;
    if (V == 0) return __ldg(reinterpret_cast<const float4*>(p));
    if (V == 1) asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    if (V == 2) asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    if (V == 3) asm volatile("ld.global.cv.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    if (V == 4) asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    if (V == 5) asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    if (V == 6) asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    if (V == 7) asm volatile("ld.global.lu.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
template <int V>
__global__ void scan_var(const float* __restrict__ base, long long n_rows, unsigned long long* out)
{
    constexpr int K = 4;
    const int lane = threadIdx.x & 31;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long n_blocks = n_rows >> 5;
    float4 q[K];
    #pragma unroll
    for (int k = 0; k < K; ++k) {
        const long long blk = gw + k * nw;
        q[k] = make_float4(0, 0, 0, 0);
        if (blk < n_blocks) q[k] = ld16<V>(base + (blk * 32 + lane) * D + 4);
    }
    unsigned cnt = 0;
    for (long long blk0 = gw; blk0 < n_blocks; blk0 += K * nw) {
        #pragma unroll
        for (int k = 0; k < K; ++k) {
            const long long blk = blk0 + k * nw;
            if (blk >= n_blocks) break;
            const float4 h = q[k];
            const long long nb = blk + K * nw;
            if (nb < n_blocks) q[k] = ld16<V>(base + (nb * 32 + lane) * D + 4);
            cnt += (h.x + h.y + h.z + h.w) > 1.0f;
        }
    }
    atomicAdd(out + 1, (unsigned long long)cnt);
}

// scan by 16-byte bulk copies (one per row, issued by the row's lane) into a shared ring
__global__ void scan_bulk16(const float* __restrict__ base, long long n_rows, unsigned long long* out)
{
    constexpr int K = 8;
    __shared__ __align__(16) float4 ring[8][K][32];
    __shared__ uint64_t bars[8][K];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long n_blocks = n_rows >> 5;
    if (lane < K) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[warp][lane])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    auto issue = [&](long long blk, int st) {
        if (blk >= n_blocks) return;
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[warp][st])), "r"(512) : "memory");
        __syncwarp();
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(&ring[warp][st][lane])), "l"(base + (blk * 32 + lane) * D + 4), "r"(16),
                       "r"(smem_u32(&bars[warp][st])) : "memory");
    };
    for (int k = 0; k < K; ++k) issue(gw + k * nw, k);
    unsigned cnt = 0, phases = 0;
    int st = 0;
    for (long long blk = gw; blk < n_blocks; blk += nw) {
        unsigned done;
        do {
            asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                         : "=r"(done) : "r"(smem_u32(&bars[warp][st])), "r"((phases >> st) & 1u) : "memory");
        } while (!done);
        phases ^= 1u << st;
        const float4 h = ring[warp][st][lane];
        __syncwarp();
        issue(blk + (long long)K * nw, st);
        st = st + 1 == K ? 0 : st + 1;
        cnt += (h.x + h.y + h.z + h.w) > 1.0f;
    }
    atomicAdd(out + 1, (unsigned long long)cnt);
}

// ---- scan, K loads in flight per lane in registers --------------------------------
template <int K>
__global__ void scan_ldg(const float* __restrict__ base, long long n_rows, unsigned long long* out)
{
    const int lane = threadIdx.x & 31;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long n_blocks = n_rows >> 5;
    float4 q[K];
    #pragma unroll
    for (int k = 0; k < K; ++k) {
        const long long blk = gw + k * nw;
        q[k] = make_float4(0, 0, 0, 0);
        if (blk < n_blocks) q[k] = __ldg(reinterpret_cast<const float4*>(base + (blk * 32 + lane) * D + 4));
    }
    unsigned cnt = 0;
    for (long long blk0 = gw; blk0 < n_blocks; blk0 += K * nw) {
        #pragma unroll
        for (int k = 0; k < K; ++k) {
            const long long blk = blk0 + k * nw;
            if (blk >= n_blocks) break;
            const float4 h = q[k];
            const long long nb = blk + K * nw;
            if (nb < n_blocks) q[k] = __ldg(reinterpret_cast<const float4*>(base + (nb * 32 + lane) * D + 4));
            cnt += (h.x + h.y + h.z + h.w) > 1.0f;
        }
    }
    if (cnt == 0xffffffffu) out[0] = cnt;
    atomicAdd(out + 1, (unsigned long long)cnt);
}

// ---- scan through cp.async (LDGSTS) into a K-deep shared ring ------------------------
template <int K>
__global__ void scan_cpasync(const float* __restrict__ base, long long n_rows, unsigned long long* out)
{
    extern __shared__ float4 ring[];                   // [warps][K][32]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long n_blocks = n_rows >> 5;
    float4* my = ring + (size_t)warp * K * 32;
    auto issue = [&](long long blk, int st) {
        if (blk < n_blocks) {
            const float* src = base + (blk * 32 + lane) * D + 4;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(my + st * 32 + lane)), "l"(src));
        }
        asm volatile("cp.async.commit_group;");
    };
    #pragma unroll
    for (int k = 0; k < K; ++k) issue(gw + k * nw, k);
    unsigned cnt = 0;
    int st = 0;
    for (long long blk = gw; blk < n_blocks; blk += nw) {
        asm volatile("cp.async.wait_group %0;" ::"n"(K - 1));
        const float4 h = my[st * 32 + lane];
        issue(blk + (long long)K * nw, st);
        st = st + 1 == K ? 0 : st + 1;
        cnt += (h.x + h.y + h.z + h.w) > 1.0f;
    }
    asm volatile("cp.async.wait_group 0;");
    atomicAdd(out + 1, (unsigned long long)cnt);
}

// ---- rows: a warp copies the selected rows of its blocks into shared memory -------------
// mode 0: cp.async.bulk (one per row, issued by the row's lane), dst stride `dstride` bytes
// mode 1: cp.async 16 B x 22 lanes per row
// mode 2: cp.async.bulk with the destination at the source's offset within a 128-byte line
__global__ void rows_copy(const float* __restrict__ base, long long n_rows, unsigned pass_per_1024,
                          int mode, int dstride, unsigned long long* out)
{
    extern __shared__ __align__(128) unsigned char pool[];   // [warps][32 rows][dstride]
    __shared__ uint64_t bars[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long n_blocks = n_rows >> 5;
    unsigned char* my = pool + (size_t)warp * 32 * dstride;
    uint64_t* bar = &bars[warp];
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    unsigned phase = 0, cnt = 0;
    for (long long blk = gw; blk < n_blocks; blk += nw) {
        const long long row = blk * 32 + lane;
        unsigned h = (unsigned)(row * 2654435761u);
        h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        const bool pass = (h & 1023u) < pass_per_1024;
        const unsigned m = __ballot_sync(0xffffffffu, pass);
        const int n = __popc(m);
        if (!n) continue;
        const int rank = __popc(m & ((1u << lane) - 1u));
        if (mode == 0 || mode == 2) {
            if (lane == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n * D * 4) : "memory");
            __syncwarp();
            if (pass)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(my + rank * dstride + (mode == 2 ? (int)((row * D * 4) & 127) : 0))), "l"(base + row * D),
                               "r"(D * 4), "r"(smem_u32(bar)) : "memory");
            unsigned done;
            do {
                asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                             : "=r"(done) : "r"(smem_u32(bar)), "r"(phase) : "memory");
            } while (!done);
            phase ^= 1u;
        } else {
            unsigned rest = m;
            int slot = 0;
            while (rest) {
                const int src_lane = __ffs((int)rest) - 1;
                rest &= rest - 1;
                const long long r = blk * 32 + src_lane;
                if (lane < D / 4)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
                                 ::"r"(smem_u32(my + slot * dstride + lane * 16)), "l"(base + r * D + lane * 4));
                ++slot;
            }
            asm volatile("cp.async.commit_group;");
            asm volatile("cp.async.wait_group 0;");
            __syncwarp();
        }
        cnt += reinterpret_cast<const float*>(my)[lane] > 1.0f;
    }
    atomicAdd(out + 1, (unsigned long long)cnt);
}

template <typename F> float time_ms(F f, int reps = 5)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main(int argc, char** argv)
{
    const int only = argc > 1 ? atoi(argv[1]) : -1;          // run a single variant (for ncu)
    if (getenv("L2_FETCH")) {
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(getenv("L2_FETCH")));
        size_t v = 0; cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity);
        printf("cudaLimitMaxL2FetchGranularity: %s, now %zu\n", cudaGetErrorString(e), v);
    }
    const long long n_rows = 4096ll * 7581;
    float* d; unsigned long long* out;
    cudaMalloc(&d, n_rows * D * 4); cudaMalloc(&out, 64);
    cudaMemset(d, 0, n_rows * D * 4); cudaMemset(out, 0, 64);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const double rows_g = n_rows / 1e9;
    int id = 0;
    auto report = [&](const char* name, int a, int b, float ms) {
        printf("%2d %-14s %3d %3d  %.3f ms  %.1f Grows/s\n", id, name, a, b, ms, rows_g / (ms * 1e-3));
        fflush(stdout);
    };
    for (int wps : {6, 12, 16, 32, 64}) {
        const int ctas = sms * (wps >= 16 ? wps / 8 : 1), thr = wps >= 16 ? 256 : wps * 32;
        if (only < 0 || only == id) report("ldg K=4", wps, 4, time_ms([&] { scan_ldg<4><<<ctas, thr>>>(d, n_rows, out); })); ++id;
        if (only < 0 || only == id) report("ldg K=8", wps, 8, time_ms([&] { scan_ldg<8><<<ctas, thr>>>(d, n_rows, out); })); ++id;
    }
    for (int wps : {4, 6, 12, 32}) {
        const int ctas = sms * (wps >= 16 ? wps / 8 : 1), thr = wps >= 16 ? 256 : wps * 32;
        const int wpc = thr / 32;
        cudaFuncSetAttribute(scan_cpasync<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        cudaFuncSetAttribute(scan_cpasync<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (only < 0 || only == id) report("cpasync K=16", wps, 16, time_ms([&] { scan_cpasync<16><<<ctas, thr, wpc * 16 * 512>>>(d, n_rows, out); })); ++id;
        if (wps <= 12) { if (only < 0 || only == id) report("cpasync K=32", wps, 32, time_ms([&] { scan_cpasync<32><<<ctas, thr, wpc * 32 * 512>>>(d, n_rows, out); })); } ++id;
    }
    // rows: 18% of the rows, 16 warps per SM (2 CTAs x 8 warps), private 32-row pool per warp
    cudaFuncSetAttribute(rows_copy, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 32 * 512);
    for (int mode = 0; mode < 2; ++mode)
        for (int dstride : {352, 368, 512}) {
            if (only < 0 || only == id)
                report(mode ? "rows cp.async" : "rows bulk", mode, dstride,
                       time_ms([&] { rows_copy<<<sms * 2, 256, 8 * 32 * dstride>>>(d, n_rows, 184, mode, dstride, out); }));
            ++id;
        }
    {
        const int ctas = sms, thr = 12 * 32;
        if (only < 0 || only == id) report("var nc(ldg)", 12, 0, time_ms([&] { scan_var<0><<<ctas, thr>>>(d, n_rows, out); })); ++id;
        if (only < 0 || only == id) report("var ld.global", 12, 1, time_ms([&] { scan_var<1><<<ctas, thr>>>(d, n_rows, out); })); ++id;
        if (only < 0 || only == id) report("var .cg", 12, 2, time_ms([&] { scan_var<2><<<ctas, thr>>>(d, n_rows, out); })); ++id;
        if (only < 0 || only == id) report("var .cv", 12, 3, time_ms([&] { scan_var<3><<<ctas, thr>>>(d, n_rows, out); })); ++id;
        if (only < 0 || only == id) report("var L1noalloc", 12, 4, time_ms([&] { scan_var<4><<<ctas, thr>>>(d, n_rows, out); })); ++id;
        if (only < 0 || only == id) report("var nc.noalloc", 12, 5, time_ms([&] { scan_var<5><<<ctas, thr>>>(d, n_rows, out); })); ++id;
        if (only < 0 || only == id) report("var .cs", 12, 6, time_ms([&] { scan_var<6><<<ctas, thr>>>(d, n_rows, out); })); ++id;
        if (only < 0 || only == id) report("var .lu", 12, 7, time_ms([&] { scan_var<7><<<ctas, thr>>>(d, n_rows, out); })); ++id;
        if (only < 0 || only == id) report("scan bulk16", 8, 8, time_ms([&] { scan_bulk16<<<ctas, 256>>>(d, n_rows, out); })); ++id;
    }
    if (only < 0 || only == id)
        report("rows bulk algn", 2, 512, time_ms([&] { rows_copy<<<sms * 2, 256, 8 * 32 * 512>>>(d, n_rows, 184, 2, 512, out); }));
    ++id;
    cudaDeviceSynchronize();
    printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
