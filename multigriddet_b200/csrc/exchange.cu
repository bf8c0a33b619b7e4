// Detection exchange: the one exchange step of the image-sharded path (SURVEY.md 8e --
// evaluator.py:254-289 collects the per-image detections of a batch on one host; with the
// batch sharded over GPUs every rank needs every image's padded detection rows).
//
// The NMS kernels write each image's rows into the exchange buffers of ALL ranks as they
// produce them (nms.cu: mirror_store -- peer stores through CUDA-IPC mappings, NVLink on a
// B200 box), so there is no separate collective and no packing pass: what remains is making
// the writes visible in order, which is this file -- a one-warp barrier kernel over flag
// words in the same buffers.
//   entry barrier ("ready"):   before a rank's first remote store of a call, every peer has
//                              enqueued past the previous call's consumers (nobody still
//                              reads the rows about to be overwritten); signalled at the start
//                              of the call, waited for just before the NMS;
//   exit barrier ("complete"): after a rank's last remote store; when it returns on the
//                              stream, every rank's rows of this call are in local memory.
// Flags are monotonically increasing call counters (epochs), one word per (kind, source
// rank), written by the source with st.release.sys after a system fence and polled by the
// owner with ld.acquire.sys.  A peer that never arrives is reported, not waited for forever:
// after kTimeoutNs the kernel gives up and raises the buffer's timeout word.
#include "common.cuh"

namespace {

constexpr unsigned long long kTimeoutNs = 20ull * 1000 * 1000 * 1000;   // 20 s

__device__ __forceinline__ unsigned long long global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// phase bit 0: signal (tell every peer this rank has reached the point), bit 1: wait (for every
// peer's signal).  The "ready" barrier is issued as two launches -- the signal at the very start
// of a call, the wait just before the first remote store -- so a peer only has to have ENTERED
// its call, not finished its decoder, for this rank's NMS to start.
__global__ void __launch_bounds__(32)
exchange_barrier_kernel(const __grid_constant__ ExchangeView v, int which, unsigned epoch, int phase)
{
    const int r = threadIdx.x;
    if (r >= v.world || r == v.rank) return;
    if (phase & 1) {
        // everything this stream did before (the NMS kernels' remote stores completed with their
        // grid) is ordered before the flag by the fence's cumulativity
        __threadfence_system();
        unsigned* theirs = v.header[r] + which * MGD_EXCHANGE_MAX_RANKS + v.rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
    }
    if (!(phase & 2)) return;
    const unsigned* mine = v.header[v.rank] + which * MGD_EXCHANGE_MAX_RANKS + r;
    const unsigned long long t0 = global_ns();
    for (;;) {
        unsigned got;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(got) : "l"(mine) : "memory");
        if ((int)(got - epoch) >= 0) break;                    // (epochs wrap; compare by distance)
        if (global_ns() - t0 > kTimeoutNs) {
            atomicAdd(v.header[v.rank] + MGD_EXCHANGE_TIMEOUT_WORD, 1u);
            break;
        }
        __nanosleep(100);
    }
}

}  // namespace

cudaError_t launch_exchange_barrier(const ExchangeView& v, int which, unsigned epoch, int phase,
                                    cudaStream_t stream)
{
    if (v.world <= 1) return cudaSuccess;
    exchange_barrier_kernel<<<1, 32, 0, stream>>>(v, which, epoch, phase);
    return cudaGetLastError();
}
