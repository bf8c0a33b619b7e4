// Dense head decode + score threshold + candidate compaction for sm_100a.
//
// Replaces MultiGridDecoder.decode_predictions and the threshold step of
// handle_predictions (reference multigriddet/postprocess/multigrid_decode.py:
// 100-183, 262-278) without ever materialising the reference's (B, 7581, 85)
// float64 tensor.  Box reconstruction of the few candidates (:151-163, 185-235)
// happens in the NMS kernel, one thread per candidate.
//
// decode_ws_kernel (the hot path: 3 anchors per layer, C <= 128) -- the three filter
// levels described below, split over WARP-SPECIALISED roles.  A CTA holds groups of one
// producer warp and three consumer warps around a ring of shared-memory slots of 32 rows:
//   producer  scans the head tensor (level 1, a 16-byte load per row, several blocks in
//             flight), and has the TMA copy each surviving 352-byte row into the slot it
//             is filling (cp.async.bulk -> the slot's `full` mbarrier); a full slot is
//             published with one arrive;
//   consumers take published slots in order (shared counter), run level 2 (lane per
//             row) and level 3 (eight lanes per row, exact) straight from shared memory,
//             append the slot's candidates to the per-image lists with one atomicAdd per
//             image, and hand the slot back through its `empty` mbarrier.
// The scan's DRAM latency and the exact evaluation's arithmetic latency thus overlap
// instead of alternating inside one warp (round 1: 46% issue utilisation, 23% of the
// warp slots).  decode_compact_kernel below is the generic fallback (any anchor count,
// any channel count) with the same three levels inside every warp.
//
// decode_compact_kernel -- persistent CTAs of independent warps.  A warp walks the
// head tensor in blocks of 32 cell rows and filters them in three levels; only the
// bytes a level needs are ever requested from HBM:
//   1. lane per row, one 16-byte load (objectness + anchor logits): an upper bound
//      of the score.  score <= sigmoid(obj) * max_anchor_prob because the class
//      maximum is <= 1, so a row below confidence*(1-1e-3) can never be a candidate
//      (fast intrinsics, relative error ~1e-6 << the 1e-3 margin).  On a trained
//      head ~85% of the rows end here having cost 32 of their 352 bytes;
//   2. the surviving rows are appended to the warp's pool in shared memory with
//      one TMA bulk copy each (cp.async.bulk + mbarrier, whole 352-byte rows, pool
//      stride padded to 23 float4 so lane-per-row 16-byte reads are conflict-free).
//      When the pool fills, all 32 lanes evaluate the same bound including the
//      class softmax, one row per lane;
//   3. rows still alive are evaluated exactly, eight lanes per row, in the
//      reference's float32 operation order: softmax as exp(x - max) / sum with
//      NumPy's pairwise 8-accumulator summation (the eight lanes ARE the eight
//      accumulators), glibc-equivalent expf (libm_emul.h), first-maximum argmax on
//      the probabilities, score = (obj * anchor) * class, threshold in float64 like
//      `score >= confidence`.  Candidates are appended to the per-image list (one
//      atomic each) as 32-byte records: score, cell, class, anchor, raw box logits.
// Algorithmic bytes (what the reference reads) are cells*D*4 per image; the DRAM
// traffic of this kernel is lower and data-dependent.  No CTA barrier, no
// inter-warp dependency.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "decode_math.cuh"

namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr int kDenseThreads = 256;

// ---- PTX: mbarrier + TMA bulk copy -------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    uint32_t done;
    do {
        asm volatile(
            "{\n .reg .pred p;\n"
            " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            " selp.u32 %0, 1, 0, p;\n}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}

struct WarpPool {
    float* rows;          // [kPoolRows][stride] floats
    int* cell;            // [kPoolRows] row index within the layer
    float* bound;         // [kPoolRows] level-1 bound
    uint64_t* bar;        // bulk-copy completion
    int stride;           // floats per pool row (16-byte multiple when TMA is used)
};

// Level 2 + 3 on the `count` rows currently in the warp's pool (all of one layer).
// Called by all 32 lanes.
template <bool kFast, int kPoolRows>
__device__ __noinline__ void flush_pool(const DecodeArgs& a, const WarpPool& pool, int count,
                                        int layer, const uint64_t* s_tab)
{
    const HeadGeom& g = a.g;
    const int lane = threadIdx.x & 31;
    const int A = g.na[layer], C = g.C;
    const bool softmax = a.use_softmax != 0;
    const bool rescore = a.rescore != 0;

    // ---- level 2: lane per pooled row, bound including the class maximum --------------
    bool pass = lane < count;
    if (pass && rescore) {
        const float* c = pool.rows + lane * pool.stride + 5 + A;
        float mc = -INFINITY, sum = 0.f;
        if (kFast) {                            // A == 3: classes start at float 8, C % 4 == 0
            const float4* c4 = reinterpret_cast<const float4*>(c);
            #pragma unroll 4
            for (int i = 0; i < C / 4; ++i) {
                const float4 v = c4[i];
                mc = fmaxf(fmaxf(mc, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
                if (softmax) sum += (fast_exp(v.x) + fast_exp(v.y)) + (fast_exp(v.z) + fast_exp(v.w));
            }
        } else {
            for (int i = 0; i < C; ++i) {
                mc = fmaxf(mc, c[i]);
                if (softmax) sum += fast_exp(c[i]);
            }
        }
        float ub;
        if (softmax) {
            // max softmax = exp(mc) / sum; no max subtraction in the fast bound: if it
            // overflows (logits > 88) the row is simply kept for the exact evaluation
            ub = __fdividef(pool.bound[lane] * fast_exp(mc), sum);
            if (!(sum < 3.0e38f) || !(ub == ub)) ub = 3.0e38f;
        } else {
            ub = pool.bound[lane] * fast_sigmoid(mc);
        }
        pass = ub >= a.score_lo;
    }
    unsigned todo = __ballot_sync(0xffffffffu, pass);

    // ---- level 3: exact evaluation, octet per row --------------------------------------
    const int j = lane & 7;
    const int oct = lane >> 3;
    const unsigned omask = 0xffu << (oct * 8);
    const int cells_l = g.gh[layer] * g.gw[layer];
    while (todo) {
        const unsigned pos = nth_set_bit(todo, oct);
        if (pos < 32u) {
            float* x = pool.rows + pos * pool.stride;
            float pa, pc; int ka, kc;
            octet_probs(x + 5, A, j, omask, softmax, s_tab, pa, ka);
            octet_probs(x + 5 + A, C, j, omask, softmax, s_tab, pc, kc);
            float score = mgd_expitf_tab(x[4], s_tab);                       // :147
            if (rescore) score = __fmul_rn(__fmul_rn(score, pa), pc);        // :170
            if (j == 0 && (double)score >= a.confidence) {                   // :271
                const int grow = pool.cell[pos];
                const int b = grow / cells_l;
                const int cell = grow - b * cells_l;
                Cand cd;
                cd.score = score;
                cd.index = g.cell_off[layer] + cell;
                cd.t[0] = x[0]; cd.t[1] = x[1]; cd.t[2] = x[2]; cd.t[3] = x[3];
                cd.cls = kc;
                cd.anchor = g.anchor_first[layer] + ka;
                const int slot = atomicAdd(a.counts + b, 1);
                a.cand[(size_t)b * g.cells + slot] = cd;
            }
        }
        __syncwarp();
        #pragma unroll
        for (int q = 0; q < 4; ++q) todo &= todo - 1;
    }
    // pool rows were rewritten in place through the generic proxy; order those writes
    // before the bulk copies (async proxy) that will refill the slots
    if (kFast) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
}

// kFast: every layer has A == 3 anchors, D % 4 == 0 and 16-byte aligned tensors:
// 16-byte loads for level 1, TMA bulk copies into the pool, float4 pool reads.
template <bool kFast, int kPoolRows>
__global__ void __launch_bounds__(kThreads, kPoolRows >= 32 ? 2 : (kPoolRows >= 24 ? 3 : (kPoolRows >= 16 ? 4 : 2)))
decode_compact_kernel(const __grid_constant__ DecodeArgs a, int pool_stride)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t s_bar[kWarpsPerCta];
    __shared__ uint64_t s_tab[MGD_EXP2F_N];
    __shared__ int s_cell[kWarpsPerCta][kPoolRows];
    __shared__ float s_bound[kWarpsPerCta][kPoolRows];

    const HeadGeom& g = a.g;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    WarpPool pool;
    pool.rows = reinterpret_cast<float*>(smem_raw) + (size_t)warp * kPoolRows * pool_stride;
    pool.cell = s_cell[warp];
    pool.bound = s_bound[warp];
    pool.bar = &s_bar[warp];
    pool.stride = pool_stride;

    if (tid < MGD_EXP2F_N) s_tab[tid] = mgd_exp2f_tab[tid];
    if (kFast && lane == 0) {
        mbar_init(pool.bar, 1);
        fence_mbar_init();
    }
    __syncthreads();

    const int gwarp = blockIdx.x * kWarpsPerCta + warp;
    const int n_warps = gridDim.x * kWarpsPerCta;
    const bool softmax = a.use_softmax != 0;
    const bool rescore = a.rescore != 0;
    uint32_t bar_phase = 0;

    for (int layer = 0; layer < g.L; ++layer) {
        const int D = g.D[layer];
        const int A = g.na[layer];
        const int n_rows = (int)a.rows_in_layer[layer];
        const int n_blocks = (n_rows + 31) >> 5;
        const float* base = a.pred[layer];
        int count = 0;                              // rows in the pool (all of this layer)

        // level-1 inputs of a block: [obj, anchor logits]; prefetched one block ahead
        auto load_head = [&](int blk, float4& q) {
            const int row = blk * 32 + lane;
            q = make_float4(NAN, 0.f, 0.f, 0.f);          // NaN never passes a >= test
            if (blk < n_blocks && row < n_rows) {
                if (kFast) q = __ldg(reinterpret_cast<const float4*>(base + (size_t)row * D + 4));
                else q.x = __ldg(base + (size_t)row * D + 4);
            }
        };
        float4 q_next;
        load_head(gwarp, q_next);
        for (int blk = gwarp; blk < n_blocks; blk += n_warps) {
            const float4 q = q_next;
            load_head(blk + n_warps, q_next);
            const int row = blk * 32 + lane;

            // ---- level 1: lane per row ---------------------------------------------------
            float bound = 0.f;
            bool pass = false;
            if (q.x >= a.obj_logit_min) {           // rows beyond the layer carry NaN
                bound = fast_sigmoid(q.x);
                if (rescore) {
                    if (kFast) {
                        const float ma = fmaxf(q.y, fmaxf(q.z, q.w));
                        if (softmax)
                            bound = __fdividef(bound, (fast_exp(q.y - ma) + fast_exp(q.z - ma)) + fast_exp(q.w - ma));
                        else
                            bound *= fast_sigmoid(ma);
                    } else {
                        const float* an = base + (size_t)row * D + 5;
                        float ma = __ldg(an);
                        for (int i = 1; i < A; ++i) ma = fmaxf(ma, __ldg(an + i));
                        if (softmax) {
                            float sa = 0.f;
                            for (int i = 0; i < A; ++i) sa += fast_exp(__ldg(an + i) - ma);
                            bound = __fdividef(bound, sa);
                        } else {
                            bound *= fast_sigmoid(ma);
                        }
                    }
                }
                pass = bound >= a.score_lo;
            }
            unsigned todo = __ballot_sync(0xffffffffu, pass);
            if (!todo) continue;

            // ---- append the survivors' rows to the pool; flush whenever it is full ----------
            while (todo) {
                const int free_slots = kPoolRows - count;
                const int rank = __popc(todo & ((1u << lane) - 1u));
                const bool take = ((todo >> lane) & 1u) && rank < free_slots;
                const unsigned taken = __ballot_sync(0xffffffffu, take);
                const int n_new = __popc(taken);
                if (take) {
                    pool.cell[count + rank] = row;
                    pool.bound[count + rank] = bound;
                }
                if (kFast) {
                    // one bulk copy per surviving row, issued by the row's own lane; the copies
                    // stay in flight while the warp moves on -- they are awaited at the flush
                    if (lane == 0) mbar_expect_tx(pool.bar, (uint32_t)n_new * D * sizeof(float));
                    __syncwarp();
                    if (take)
                        bulk_g2s(pool.rows + (size_t)(count + rank) * pool.stride,
                                 base + (size_t)row * D, (uint32_t)D * sizeof(float), pool.bar);
                } else {
                    unsigned rest = taken;
                    int slot = count;
                    while (rest) {
                        const int src_lane = __ffs((int)rest) - 1;
                        rest &= rest - 1;
                        const float* src = base + ((size_t)blk * 32 + src_lane) * D;
                        float* dst = pool.rows + (size_t)slot * pool.stride;
                        for (int i = lane; i < D; i += 32) dst[i] = __ldg(src + i);
                        ++slot;
                    }
                    __syncwarp();
                }
                count += n_new;
                todo &= ~taken;
                if (count == kPoolRows) {
                    if (kFast) {
                        if (lane == 0) mbar_arrive(pool.bar);
                        mbar_wait(pool.bar, bar_phase);
                        bar_phase ^= 1u;
                    }
                    flush_pool<kFast, kPoolRows>(a, pool, count, layer, s_tab);
                    count = 0;
                }
            }
        }
        if (count) {
            if (kFast) {
                if (lane == 0) mbar_arrive(pool.bar);
                mbar_wait(pool.bar, bar_phase);
                bar_phase ^= 1u;
            }
            flush_pool<kFast, kPoolRows>(a, pool, count, layer, s_tab);
        }
        __syncwarp();
    }
}

// =================================================================================
// Warp-specialised decoder
// =================================================================================
namespace ws {

// Shape of a CTA: kProducers producer warps and kConsumers consumer warps around ONE ring of
// kSlots slots.  Slots are handed out by ticket: a producer takes the next sequence number when
// it has a first row to store, a consumer takes the next sequence number when it is free; both
// counters run through the same sequence, so the consumer of ticket s waits for the producer of
// ticket s, whoever that is (slot = s mod kSlots, mbarrier phase = s div kSlots).
// Measured per COCO image (planted head): producers ~43 k warp-instructions (scan + one TMA issue
// per surviving row), consumers ~80 k, both bound by dependent-issue latency: about two
// consumers per producer, and as many slots as shared memory holds (a slot is busy for the fill,
// the DRAM latency of its last row, and the exact evaluation).  Slots of 16 rows instead of 32
// halve what a slot holds while it is being filled or waiting for its last row (level 2 then
// runs two lanes per row); rows arrive at ~1.3 slots of 32 per microsecond and SM.
template <int kProducers_, int kConsumers_, int kSlots_, int kRows_ = 32, int kPrefetch_ = 16>
struct Shape {
    static constexpr int kPrefetch = kPrefetch_;                 // level-1 blocks in flight per producer (512 B each)
    static constexpr int kProducers = kProducers_;
    static constexpr int kConsumers = kConsumers_;
    static constexpr int kThreads = (kProducers_ + kConsumers_) * 32;
    static constexpr int kSlots = kSlots_;
    static constexpr int kRows = kRows_;                         // rows per slot: 32 or 16
    static constexpr int kCtasPerSm = kSlots_ * kRows_ > 9 * 32 ? 1 : 2;   // by shared memory (352 B per COCO row)
};
constexpr int kSlotRows = 32;                 // capacity of the per-slot metadata arrays

struct SlotMeta {
    int row[kSlotRows];                       // row index within the layer (b * cells_l + cell)
    float bound[kSlotRows];                   // level-1 bound sigmoid(obj) * max anchor prob
    int count;                                // rows in the slot; < 0: no more work
    int layer;
    int pad[2];
};

// exp2f-style kernel of libm_emul.h with the table split into two 32-entry 4-byte arrays
// (low words at tab[0..31], high words at tab[32..63]; one entry per bank: any lane pattern
// is conflict-free) and the 64-bit exponent add done
// on the high word only (ki << 47 has no low half).  Same double operations, same bits.
__device__ __forceinline__ float expf_core2(float x, const uint32_t* tab)
{
    const double inv_ln2_n = 0x1.71547652b82fep+0 * MGD_EXP2F_N;
    const double shift = 0x1.8p+52;
    const double c0 = 0x1.c6af84b912394p-5 / MGD_EXP2F_N / MGD_EXP2F_N / MGD_EXP2F_N;
    const double c1 = 0x1.ebfce50fac4f3p-3 / MGD_EXP2F_N / MGD_EXP2F_N;
    const double c2 = 0x1.62e42ff0c52d6p-1 / MGD_EXP2F_N;
    const double z = __dmul_rn(inv_ln2_n, (double)x);
    double kd = __dadd_rn(z, shift);
    const int ki = __double2loint(kd);
    kd = __dsub_rn(kd, shift);
    const double r = __fma_rn(inv_ln2_n, (double)x, -kd);      // exact residual (libm_emul.h)
    const int idx = ki & (MGD_EXP2F_N - 1);
    const int hi = (int)tab[idx + MGD_EXP2F_N] + (ki << 15);
    const double sc = __hiloint2double(hi, (int)tab[idx]);
    const double q = __fma_rn(c0, r, c1);
    const double r2 = __dmul_rn(r, r);
    double y = __fma_rn(c2, r, 1.0);
    y = __fma_rn(q, r2, y);
    y = __dmul_rn(y, sc);
    return __double2float_rn(y);
}

// full-range version (glibc's special-case block in front, libm_emul.h: mgd_expf_tab)
__device__ __forceinline__ float expf_full2(float x, const uint32_t* tab)
{
    float special = 0.f;
    bool is_special = false;
    if (!(x < 88.0f && x > -88.0f)) {
        if (x != x) { special = x + x; is_special = true; }
        else if (x > 88.72283172607421875f) { special = __builtin_huge_valf(); is_special = true; }
        else if (x < -103.972076416015625f) { special = 0.0f; is_special = true; }
    }
    const float v = expf_core2(is_special ? 0.f : x, tab);
    return is_special ? special : v;
}

// if (d >= thr) { first = min(first, i); ++cnt; } as one compare and two predicated ops
__device__ __forceinline__ void near_track(float d, float thr, int i, int& first, int& cnt)
{
    asm("{\n .reg .pred p;\n setp.ge.f32 p, %2, %3;\n @p min.s32 %0, %0, %4;\n @p add.s32 %1, %1, 1;\n}"
        : "+r"(first), "+r"(cnt) : "f"(d), "f"(thr), "r"(i));
}

__device__ __forceinline__ int ld_acquire_smem(const int* p)
{
    int v;
    asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_smem(int* p, int v)
{
    asm volatile("st.release.cta.shared.s32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}

struct GroupShared {
    float4* stage;            // [kPrefetch][32]: level-1 inputs in flight (this producer's own)
    float* rows;              // [kSlots][kSlotRows][stride]
    SlotMeta* meta;           // [kSlots]
    uint64_t* full;           // [kSlots]: TMA bytes of the slot's current use have landed
    int* published;           // [kSlots]: ticket most recently published in the slot
    int* released;            // [kSlots]: completed uses of the slot
    int* fill_seq;            // producers' ticket counter
    int* next_seq;            // consumers' ticket counter
    int* producers_done;
    int stride;
};

template <class Sh>
__device__ __forceinline__ void producer(const DecodeArgs& a, const GroupShared& gs, int p, int P)
{
    constexpr int kSlots = Sh::kSlots;
    constexpr int kRows = Sh::kRows;
    constexpr int kPrefetch = Sh::kPrefetch;
    constexpr int kConsumers = Sh::kConsumers;
    constexpr int kProducers = Sh::kProducers;
    const HeadGeom& g = a.g;
    const int lane = threadIdx.x & 31;
    const bool softmax = a.use_softmax != 0;
    const bool rescore = a.rescore != 0;
    int seq = 0, slot = 0, count = 0;
    bool have_slot = false;

    // Tickets from different producers complete out of order, so a waiter can be several uses of
    // a slot ahead of the slot's state; an mbarrier only distinguishes the parity of a phase.
    // The hand-over therefore goes through per-slot sequence words (released / published, with
    // release / acquire semantics); the mbarrier is only waited on for the TMA bytes, once the
    // slot is known to be in the waiter's own use.
    auto acquire = [&]() {
        if (lane == 0) {
            seq = atomicAdd(gs.fill_seq, 1);
            const int use = seq / kSlots;
            while (ld_acquire_smem(&gs.released[seq % kSlots]) != use) __nanosleep(40);
        }
        seq = __shfl_sync(0xffffffffu, seq, 0);
        slot = seq % kSlots;
        have_slot = true;
    };
    auto publish = [&](int cnt, int layer) {
        if (lane == 0) { gs.meta[slot].count = cnt; gs.meta[slot].layer = layer; }
        __syncwarp();                                   // metadata of all lanes before the release
        if (lane == 0) {
            mbar_arrive(&gs.full[slot]);
            st_release_smem(&gs.published[slot], seq);
        }
        have_slot = false;
        count = 0;
    };

    for (int layer = 0; layer < g.L; ++layer) {
        const int D = g.D[layer];
        const int n_rows = (int)a.rows_in_layer[layer];
        const int n_blocks = (n_rows + 31) >> 5;
        const float* base = a.pred[layer];
        // Level-1 inputs ([obj, 3 anchor logits] = 16 bytes of every row) arrive through a ring
        // of cp.async copies, kPrefetch blocks ahead: the scan is DRAM-bound (the memory system
        // moves a 128-byte line per row whatever the request size) and needs far more loads in
        // flight than registers could hold.  Every lane reads back only what it copied itself.
        auto issue_head = [&](int blk, int st) {
            const int row = blk * 32 + lane;
            if (blk < n_blocks && row < n_rows)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
                             ::"r"(smem_u32(gs.stage + st * 32 + lane)), "l"(base + (size_t)row * D + 4) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        // a producer walks chunks of kChunk consecutive blocks (DRAM page locality; the rows of
        // a slot then belong to one or two images), the chunks strided over the producers;
        // kPrefetch is a multiple of the chunk, so the block kPrefetch steps ahead is kPrefetch*P on
        const int kChunk = a.chunk_blocks;              // 1, 2, 4, 8 or 16 (divides kPrefetch)
        #pragma unroll
        for (int k = 0; k < kPrefetch; ++k) issue_head((p + (k / kChunk) * P) * kChunk + k % kChunk, k);
        // survivors: one TMA copy per row into the slot being filled
        auto stage_rows = [&](int blk, unsigned todo, float bound) {
            const int row = blk * 32 + lane;
            while (todo) {
                if (!have_slot) acquire();
                const int free_slots = kRows - count;
                const int rank = __popc(todo & ((1u << lane) - 1u));
                const bool take = ((todo >> lane) & 1u) && rank < free_slots;
                const unsigned taken = __ballot_sync(0xffffffffu, take);
                const int n_new = __popc(taken);
                if (take) {
                    gs.meta[slot].row[count + rank] = row;
                    gs.meta[slot].bound[count + rank] = bound;
                }
                if (lane == 0) mbar_expect_tx(&gs.full[slot], (uint32_t)n_new * D * sizeof(float));
                __syncwarp();
                if (take)
                    bulk_g2s(gs.rows + ((size_t)slot * kRows + count + rank) * gs.stride,
                             base + (size_t)row * D, (uint32_t)D * sizeof(float), &gs.full[slot]);
                count += n_new;
                todo &= ~taken;
                if (count == kRows) publish(count, layer);
            }
        };
        int st = 0;
        for (int chunk = p; chunk * kChunk < n_blocks; chunk += P) {
            for (int u = 0; u < kChunk; ++u) {
                const int blk = chunk * kChunk + u;
                asm volatile("cp.async.wait_group %0;" ::"n"(kPrefetch - 1) : "memory");
                const int row = blk * 32 + lane;
                float4 h = gs.stage[st * 32 + lane];
                if (row >= n_rows) h.x = NAN;               // NaN never passes a >= test
                issue_head(blk + kPrefetch * P, st);
                st = st + 1 == kPrefetch ? 0 : st + 1;

                // ---- level 1: lane per row -----------------------------------------------
                float bound = 0.f;
                bool pass = false;
                if (h.x >= a.obj_logit_min) {           // rows beyond the layer carry NaN
                    bound = fast_sigmoid(h.x);
                    if (rescore) {
                        const float ma = fmaxf(h.y, fmaxf(h.z, h.w));
                        if (softmax)
                            bound = __fdividef(bound, (fast_exp(h.y - ma) + fast_exp(h.z - ma)) + fast_exp(h.w - ma));
                        else
                            bound *= fast_sigmoid(ma);
                    }
                    pass = bound >= a.score_lo;
                }
                unsigned todo = __ballot_sync(0xffffffffu, pass);
                if (a.debug & 1) todo = 0;
                stage_rows(blk, todo, bound);
            }
        }
        if (count) publish(count, layer);              // slots never mix layers
    }
    // the last producer to finish hands every consumer a terminal slot
    int done = 0;
    if (lane == 0) done = atomicAdd(gs.producers_done, 1) + 1;
    done = __shfl_sync(0xffffffffu, done, 0);
    if (done == kProducers)
        for (int c = 0; c < kConsumers; ++c) {
            acquire();
            publish(-1, 0);
        }
}

// kC: the class count when it is known at compile time (80: COCO, the benchmark), else 0.
// With a constant trip count the whole exact evaluation of a round -- ten class chains, the
// anchor / objectness chain and the three reciprocals -- is one basic block the scheduler can
// interleave; the consumers are bound by dependent-issue latency, not by issue slots.
template <class Sh, int kC>
__device__ __forceinline__ void consumer(const DecodeArgs& a, const GroupShared& gs,
                                         const uint32_t* tab)
{
    constexpr int kSlots = Sh::kSlots;
    constexpr int kRows = Sh::kRows;
    constexpr int kSplit = 32 / kRows;                  // lanes per row in level 2
    const HeadGeom& g = a.g;
    const int lane = threadIdx.x & 31;
    const int j = lane & 7;
    const int oct = lane >> 3;
    const bool softmax = a.use_softmax != 0;
    const bool rescore = a.rescore != 0;
    const int C = kC ? kC : g.C;
    const int body = C >= 8 ? (C & ~7) : 0;             // NumPy: 8 accumulators over the body,
    const int tail = C - body;                          // then the tail one by one (n < 8: all tail)
    constexpr float kNear = 0.99999f;
    constexpr float kNearD = -1.1e-5f;                  // exp(d) >= kNear implies d > kNearD

    for (;;) {
        int seq = 0;
        if (lane == 0) {
            seq = atomicAdd(gs.next_seq, 1);
            while (ld_acquire_smem(&gs.published[seq % kSlots]) != seq) __nanosleep(40);
        }
        seq = __shfl_sync(0xffffffffu, seq, 0);
        const int slot = seq % kSlots;
        mbar_wait(&gs.full[slot], ((unsigned)(seq / kSlots)) & 1u);    // the TMA bytes of this use
        const SlotMeta& meta = gs.meta[slot];
        const int count = meta.count;
        if (count < 0) break;
        const int layer = meta.layer;
        const float* rows = gs.rows + (size_t)slot * kRows * gs.stride;

        // ---- level 2: lane per row (two lanes per row with 16-row slots), bound including the
        // class maximum ------------------------------------------------------------------------
        const int lrow = lane & (kRows - 1), part = lane / kRows;
        bool pass = lrow < count && !(a.debug & 4);
        float mc = -INFINITY, sum = 0.f;
        if (pass) {
            const float4* c4 = reinterpret_cast<const float4*>(rows + (size_t)lrow * gs.stride + 8);
            const int n4 = C / 4, per = (n4 + kSplit - 1) / kSplit;
            const int i0 = part * per, i1 = min(n4, i0 + per);
            if (softmax && rescore) {
                #pragma unroll 4
                for (int i = i0; i < i1; ++i) {
                    const float4 v = c4[i];
                    mc = fmaxf(fmaxf(mc, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
                    sum += (fast_exp(v.x) + fast_exp(v.y)) + (fast_exp(v.z) + fast_exp(v.w));
                }
            } else {
                #pragma unroll 4
                for (int i = i0; i < i1; ++i) {
                    const float4 v = c4[i];
                    mc = fmaxf(fmaxf(mc, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
                }
            }
        }
        if (kSplit == 2) {
            mc = fmaxf(mc, __shfl_xor_sync(0xffffffffu, mc, 16));
            sum += __shfl_xor_sync(0xffffffffu, sum, 16);
        }
        if (pass && rescore) {
            if (softmax) {
                // max softmax = exp(mc) / sum; no max subtraction in the fast bound: if it
                // overflows (logits > 88) the row is simply kept for the exact evaluation
                float ub = __fdividef(meta.bound[lrow] * fast_exp(mc), sum);
                if (!(sum < 3.0e38f) || !(ub == ub)) ub = 3.0e38f;
                pass = ub >= a.score_lo;
            } else {
                pass = meta.bound[lrow] * fast_sigmoid(mc) >= a.score_lo;
            }
        }
        unsigned todo = __ballot_sync(0xffffffffu, pass) & (kRows == 32 ? 0xffffffffu : (1u << (kRows & 31)) - 1u);
        if (a.debug & 2) todo = 0;

        // ---- level 3: exact evaluation, octet per row ----------------------------------------
        unsigned cand_rows = 0;                         // rows of this slot that became candidates
        while (todo) {
            const unsigned pos = nth_set_bit(todo, oct);
            const bool live = pos < 32u;
            const unsigned posr = live ? pos : (unsigned)__ffs((int)todo) - 1u;   // idle octets shadow a live row
            const float* x = rows + (size_t)posr * gs.stride;
            const float mx = __shfl_sync(0xffffffffu, mc, (int)posr);
            const float4 head = *reinterpret_cast<const float4*>(x + 4);         // obj, 3 anchor logits
            const int l0 = oct * 8;
            float pa, pc, obj;
            int ka, kc;
            if (softmax) {
                // -- anchors (lanes 0-2 of the octet) and objectness (lane 3) in one expf round --
                const float ma = fmaxf(head.y, fmaxf(head.z, head.w));
                const float logit = x[5 + (j < 2 ? j : 2)];
                const float da = __fsub_rn(logit, ma);
                const float arg = j < 3 ? (da < -104.0f ? -104.0f : da) : (j == 3 ? -head.x : 0.f);
                const float ea = expf_full2(arg, tab);

                // -- classes: exp(x - max) accumulated in NumPy's order as it is produced --
                const float* cx = x + 8;
                float r = 0.f, e_tail = 0.f;            // (0 + e == e exactly: e >= 0 or NaN)
                int near_cnt = 0, near_first = INT_MAX;
                if (body) {
                    #pragma unroll 10
                    for (int i = j; i < body; i += 8) {
                        // (a select, not fmaxf: NaN -- from a NaN or +inf logit -- must reach the
                        //  sum like it does in the reference, which then drops the row)
                        const float d = __fsub_rn(cx[i], mx);
                        const float e = expf_core2(d < -104.0f ? -104.0f : d, tab);
                        r = __fadd_rn(r, e);
                        near_track(d, kNearD, i, near_first, near_cnt);
                    }
                    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));   // (r0+r1) (r2+r3) ...
                    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
                    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
                }
                if (tail) {
                    if (j < tail) {
                        const float d = __fsub_rn(cx[body + j], mx);
                        e_tail = expf_core2(d < -104.0f ? -104.0f : d, tab);
                        near_track(d, kNearD, body + j, near_first, near_cnt);
                    }
                    for (int k = 0; k < tail; ++k)
                        r = __fadd_rn(r, __shfl_sync(0xffffffffu, e_tail, l0 + k));
                }
                pc = __frcp_rn(r);                      // 1 / sum: the maximum's exponential is exactly 1
                const float e0 = __shfl_sync(0xffffffffu, ea, l0);
                const float e1 = __shfl_sync(0xffffffffu, ea, l0 + 1);
                const float e2 = __shfl_sync(0xffffffffu, ea, l0 + 2);
                const float eo = __shfl_sync(0xffffffffu, ea, l0 + 3);
                const float sa = __fadd_rn(__fadd_rn(__fadd_rn(0.f, e0), e1), e2);  // n < 8: sequential
                pa = __frcp_rn(sa);
                obj = __frcp_rn(__fadd_rn(1.0f, eo));                            // :147
                // argmax on the probabilities e/s like the reference: first index whose quotient
                // equals the maximum quotient; only exponentials within 1e-5 of 1 can tie.  A row
                // with a single logit within kNearD of the maximum (the maximum itself: a superset
                // test, exp(-1.1e-5) < 0.99999) needs no division; anything else is settled exactly
                const unsigned has = (__ballot_sync(0xffffffffu, near_cnt > 0) >> l0) & 0xffu;
                const unsigned many = (__ballot_sync(0xffffffffu, near_cnt > 1) >> l0) & 0xffu;
                const bool simple = !many && __popc(has) == 1;
                kc = __shfl_sync(0xffffffffu, near_first, l0 + (has ? __ffs((int)has) - 1 : 0));
                const int n0 = e0 >= kNear, n1 = e1 >= kNear, n2 = e2 >= kNear;
                ka = n0 ? 0 : (n1 ? 1 : 2);
                if (__any_sync(0xffffffffu, !simple || n0 + n1 + n2 != 1)) {  // rare
                    int first = INT_MAX;
                    for (int i = j; i < C; i += 8) {
                        const float d = __fsub_rn(cx[i], mx);
                        const float e = expf_core2(d < -104.0f ? -104.0f : d, tab);
                        if (e >= kNear && __fdiv_rn(e, r) == pc) { first = i; break; }
                    }
                    first = min(first, __shfl_xor_sync(0xffffffffu, first, 1));
                    first = min(first, __shfl_xor_sync(0xffffffffu, first, 2));
                    first = min(first, __shfl_xor_sync(0xffffffffu, first, 4));
                    if (!simple) kc = first;
                    if (n0 + n1 + n2 != 1) {
                        ka = INT_MAX;
                        if (n2 && __fdiv_rn(e2, sa) == pa) ka = 2;
                        if (n1 && __fdiv_rn(e1, sa) == pa) ka = 1;
                        if (n0 && __fdiv_rn(e0, sa) == pa) ka = 0;
                    }
                }
            } else {
                // element-wise expit; maximum and its first index
                float best = -INFINITY;
                int first = INT_MAX;
                for (int i = j; i < C; i += 8) {
                    const float qv = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf_full2(-x[8 + i], tab)));
                    if (qv > best) { best = qv; first = i; }
                }
                #pragma unroll
                for (int d = 1; d <= 4; d <<= 1) {
                    const float ob = __shfl_xor_sync(0xffffffffu, best, d);
                    const int oi = __shfl_xor_sync(0xffffffffu, first, d);
                    if (ob > best || (ob == best && oi < first)) { best = ob; first = oi; }
                }
                pc = best; kc = first;
                const float logit = j == 0 ? head.x : (j == 1 ? head.y : (j == 2 ? head.z : head.w));
                const float qv = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf_full2(-logit, tab)));
                obj = __shfl_sync(0xffffffffu, qv, l0);
                const float q0 = __shfl_sync(0xffffffffu, qv, l0 + 1);
                const float q1 = __shfl_sync(0xffffffffu, qv, l0 + 2);
                const float q2 = __shfl_sync(0xffffffffu, qv, l0 + 3);
                pa = -INFINITY; ka = INT_MAX;
                if (q0 > pa) { pa = q0; ka = 0; }
                if (q1 > pa) { pa = q1; ka = 1; }
                if (q2 > pa) { pa = q2; ka = 2; }
            }
            float score = obj;
            if (rescore) score = __fmul_rn(__fmul_rn(score, pa), pc);            // :170
            // A candidate leaves its result in the row's own (now dead) objectness / anchor words;
            // the records go out once per slot, below.
            const bool is_cand = live && j == 0 && (double)score >= a.confidence;   // :271
            __syncwarp();                               // shadow octets are done reading the row head
            if (is_cand) {
                float4 res;
                res.x = score;
                res.y = __int_as_float(kc);
                res.z = __int_as_float(g.anchor_first[layer] + ka);
                res.w = 0.f;
                *reinterpret_cast<float4*>(const_cast<float*>(x) + 4) = res;
            }
            cand_rows |= __reduce_or_sync(0xffffffffu, is_cand ? (1u << pos) : 0u);
            #pragma unroll
            for (int q = 0; q < 4; ++q) todo &= todo - 1;
        }
        // ---- records out: lane per row; one atomicAdd per image present in the slot ----------
        // (a per-candidate atomicAdd serialises in L2: every warp of the grid works on the same
        //  few images at any time, i.e. on the same few counters)
        __syncwarp();
        const bool mine = lane < kRows && ((cand_rows >> lane) & 1u);
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f), res = t;
        unsigned grow = 0;
        if (mine) {
            const float* x = rows + (size_t)lane * gs.stride;
            t = *reinterpret_cast<const float4*>(x);
            res = *reinterpret_cast<const float4*>(x + 4);
            grow = (unsigned)meta.row[lane];
        }
        if (cand_rows) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes before the next TMA fill
        __syncwarp();                                   // every lane is done with the slot
        if (lane == 0) st_release_smem(&gs.released[slot], seq / kSlots + 1);
        if (cand_rows) {
            const unsigned long long magic = a.cells_magic[layer];
            const unsigned b = magic ? (unsigned)__umul64hi((unsigned long long)grow, magic) : grow;
            const unsigned same = __match_any_sync(0xffffffffu, mine ? b : (0x80000000u | (unsigned)lane));
            const int leader = __ffs((int)same) - 1;
            int base = 0;
            if (mine && lane == leader) base = atomicAdd(a.counts + b, __popc(same));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (mine) {
                Cand cd;
                cd.score = res.x;
                cd.index = g.cell_off[layer] + (int)(grow - b * (unsigned)(g.gh[layer] * g.gw[layer]));
                cd.t[0] = t.x; cd.t[1] = t.y; cd.t[2] = t.z; cd.t[3] = t.w;
                cd.cls = __float_as_int(res.y);
                cd.anchor = __float_as_int(res.z);
                a.cand[(size_t)b * g.cells + base + __popc(same & ((1u << lane) - 1u))] = cd;
            }
        }
    }
}

template <class Sh, int kC>
__global__ void __launch_bounds__(Sh::kThreads, Sh::kCtasPerSm)
decode_ws_kernel(const __grid_constant__ DecodeArgs a, int stride)
{
    constexpr int kSlots = Sh::kSlots, kProducers = Sh::kProducers;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t s_full[kSlots];
    __shared__ int s_published[kSlots], s_released[kSlots];
    __shared__ SlotMeta s_meta[kSlots];
    __shared__ uint32_t s_tab[2 * MGD_EXP2F_N];
    __shared__ int s_counters[3];


    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    if (tid < MGD_EXP2F_N) {
        s_tab[tid] = (uint32_t)mgd_exp2f_tab[tid];
        s_tab[tid + MGD_EXP2F_N] = (uint32_t)(mgd_exp2f_tab[tid] >> 32);
    }
    if (tid == 0) {
        for (int sl = 0; sl < kSlots; ++sl) {
            mbar_init(&s_full[sl], 1);
            s_published[sl] = -1;
            s_released[sl] = 0;
        }
        s_counters[0] = s_counters[1] = s_counters[2] = 0;
        fence_mbar_init();
    }
    __syncthreads();

    GroupShared gs;
    gs.rows = reinterpret_cast<float*>(smem_raw);
    // producers take the HIGHEST warp ids: the issue arbiter favours them, and a producer is a
    // serial instruction stream that everything else waits for
    constexpr int kFirstProducer = Sh::kConsumers;
    const int pw = warp - kFirstProducer;
    gs.stage = reinterpret_cast<float4*>(gs.rows + (size_t)kSlots * Sh::kRows * stride) +
               (size_t)(pw >= 0 ? pw : 0) * Sh::kPrefetch * 32;
    gs.meta = s_meta;
    gs.full = s_full;
    gs.published = s_published;
    gs.released = s_released;
    gs.fill_seq = &s_counters[0];
    gs.next_seq = &s_counters[1];
    gs.producers_done = &s_counters[2];
    gs.stride = stride;
    if (pw >= 0)
        producer<Sh>(a, gs, blockIdx.x * kProducers + pw, gridDim.x * kProducers);
    else
        consumer<Sh, kC>(a, gs, s_tab);
}

}  // namespace ws

// ---- dense decode (decode_predictions API), one octet per row, not a hot path --
__global__ void __launch_bounds__(kDenseThreads, 3)
decode_dense_kernel(const __grid_constant__ DecodeArgs a, const int* image_hw, double* out,
                    int row_floats)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t s_tab[MGD_EXP2F_N];
    const HeadGeom& g = a.g;
    const int tid = threadIdx.x, j = tid & 7;
    const unsigned omask = 0xffu << (((tid & 31) >> 3) * 8);
    float* x = reinterpret_cast<float*>(smem_raw) + (size_t)(tid >> 3) * row_floats;
    if (tid < MGD_EXP2F_N) s_tab[tid] = mgd_exp2f_tab[tid];
    __syncthreads();
    const long long total_rows = (long long)a.B * g.cells;
    const long long octs = (long long)gridDim.x * (kDenseThreads / 8);
    // uniform trip count per warp: every lane iterates while ANY octet of the grid might
    const long long iters = (total_rows + octs - 1) / octs;
    for (long long it = 0; it < iters; ++it) {
        long long q = it * octs + (long long)blockIdx.x * (kDenseThreads / 8) + (tid >> 3);
        const bool live = q < total_rows;
        if (!live) q = total_rows - 1;
        const int b = (int)(q / g.cells);
        const int flat = (int)(q - (long long)b * g.cells);
        int layer = 0;
        while (layer + 1 < g.L && flat >= g.cell_off[layer + 1]) ++layer;
        const int cell = flat - g.cell_off[layer];
        const int D = g.D[layer], A = g.na[layer];
        const float* src = a.pred[layer] + ((size_t)b * g.gh[layer] * g.gw[layer] + cell) * D;
        for (int i = j; i < D; i += 8) x[i] = __ldg(src + i);
        __syncwarp(omask);
        float pa, pc; int ka, kc;
        octet_probs(x + 5, A, j, omask, a.use_softmax != 0, s_tab, pa, ka);
        octet_probs(x + 5 + A, g.C, j, omask, a.use_softmax != 0, s_tab, pc, kc);
        __syncwarp(omask);
        float score = mgd_expitf_tab(x[4], s_tab);
        if (a.rescore) score = __fmul_rn(__fmul_rn(score, pa), pc);
        const int rr = cell / g.gw[layer], cc = cell - rr * g.gw[layer];
        double box[4];
        Letterbox lb;
        if (image_hw) lb = letterbox_consts(g.in_h, g.in_w, image_hw[2 * b], image_hw[2 * b + 1]);
        decode_axis_of(g, x, layer, g.anchor_first[layer] + ka, rr, cc, image_hw ? &lb : nullptr, 0,
                       s_tab, box[0], box[2]);
        decode_axis_of(g, x, layer, g.anchor_first[layer] + ka, rr, cc, image_hw ? &lb : nullptr, 1,
                       s_tab, box[1], box[3]);
        if (live) {
            double* o = out + (size_t)q * (5 + g.C);
            if (j < 4) o[j] = box[j];
            if (j == 4) o[4] = (double)score;
            float s = 1.0f;
            if (a.use_softmax) s = np_sum_octet(x + 5 + A, g.C, j, omask);
            for (int i = j; i < g.C; i += 8) {
                const float e = x[5 + A + i];
                o[5 + i] = (double)(a.use_softmax ? __fdiv_rn(e, s) : e);
            }
        } else if (a.use_softmax) {
            (void)np_sum_octet(x + 5 + A, g.C, j, omask);
        }
        __syncwarp(omask);
    }
}

}  // namespace

cudaError_t launch_decode(const DecodeArgs& a_in, int num_sms, cudaStream_t stream)
{
    DecodeArgs a = a_in;
    const HeadGeom& g = a.g;
    int dmax = 0;
    bool fast = true;
    for (int l = 0; l < g.L; ++l) {
        dmax = g.D[l] > dmax ? g.D[l] : dmax;
        fast = fast && g.na[l] == 3 && (g.D[l] % 4 == 0) &&
               ((reinterpret_cast<uintptr_t>(a.pred[l]) & 15) == 0);
        a.rows_in_layer[l] = (long long)a.B * g.gh[l] * g.gw[l];
        if (a.rows_in_layer[l] >= 0x7fffffffll - 64) return cudaErrorInvalidValue;
        const unsigned long long cells_l = (unsigned long long)g.gh[l] * g.gw[l];
        a.cells_magic[l] = cells_l > 1 ? ~0ull / cells_l + 1ull : 0ull;
    }
    // pool row stride: an odd number of float4 (fast path) / an odd number of floats
    // (generic path) so that lane-per-row reads hit distinct banks
    int stride = fast ? (dmax / 4 | 1) * 4 : (dmax | 1);
    if (fast && stride < dmax) stride += 8;

    // ---- warp-specialised kernel: 3 anchors per layer, up to 128 classes ----------------
    static int env_impl = -1;      // MGD_DECODE_IMPL=legacy: the round-1 kernel (A/B measurements)
    static int env_shape = 0;      // MGD_DECODE_SHAPE=1|2: alternative CTA shapes (measurements)
    if (env_impl < 0) {
        const char* e = getenv("MGD_DECODE_IMPL");
        env_impl = e && !strcmp(e, "legacy") ? 1 : 0;
        const char* sh = getenv("MGD_DECODE_SHAPE");
        env_shape = sh ? atoi(sh) : 0;
    }
    static int env_debug = -1;
    if (env_debug < 0) { const char* e = getenv("MGD_DECODE_DEBUG"); env_debug = e ? atoi(e) : 0; }
    a.debug = env_debug;
    if (fast && g.C <= 128 && !env_impl) {
        long long blocks = 0;
        for (int l = 0; l < g.L; ++l) blocks += (a.rows_in_layer[l] + 31) / 32;
        // slot rows are stored back to back (stride = D floats): level 3's octet reads are
        // conflict-free at any stride, level 2's lane-per-row float4 reads take a 2-way conflict
        // at 88 floats -- cheaper than the shared memory a padded stride costs (a fourth slot)
        const int ws_stride = dmax;
        auto run = [&](auto shape) -> cudaError_t {
            using Sh = decltype(shape);
            const size_t smem = (size_t)Sh::kSlots * Sh::kRows * ws_stride * sizeof(float) +
                                (size_t)Sh::kProducers * Sh::kPrefetch * 32 * sizeof(float4);
            // (static shared memory: slot metadata, barriers, tables; 1 KB reserved per CTA)
            const size_t fixed = 1024 + 1024 + (size_t)Sh::kSlots * 304;
            int ctas_per_sm = (int)((227 * 1024) / (smem + fixed));
            if (ctas_per_sm > Sh::kCtasPerSm) ctas_per_sm = Sh::kCtasPerSm;
            if (ctas_per_sm < 1) return cudaErrorInvalidConfiguration;
            long long grid = (long long)num_sms * ctas_per_sm;
            // a producer should have a few blocks of every layer to walk
            const long long needed = (blocks + Sh::kProducers * 4 - 1) / (Sh::kProducers * 4);
            if (grid > needed) grid = needed;
            if (grid < 1) grid = 1;
            // big batches: chunks of 8 consecutive blocks per producer (DRAM page locality); mid-sized
            // ones pairs (the tail imbalance of 8-block chunks costs 256 images 18 %: 103 -> 84 us,
            // 1 024 images 2 %); small ones single blocks (latency: every producer should get work)
            const long long per_producer = blocks / (grid * Sh::kProducers);
            a.chunk_blocks = per_producer >= 768 ? 8 : (per_producer >= 64 ? 2 : 1);
            static int env_cb = -1;
            if (env_cb < 0) { const char* e = getenv("MGD_DECODE_CHUNK_BLOCKS"); env_cb = e ? atoi(e) : 0; }
            if (env_cb == 1 || env_cb == 2 || env_cb == 4 || env_cb == 8) a.chunk_blocks = env_cb;

            auto kernel = g.C == 80 ? ws::decode_ws_kernel<Sh, 80> : ws::decode_ws_kernel<Sh, 0>;
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            prof_mark_begin(PROF_DECODE_COMPACT, stream);
            kernel<<<(unsigned)grid, Sh::kThreads, smem, stream>>>(a, ws_stride);
            prof_mark_end(PROF_DECODE_COMPACT, stream);
            return cudaGetLastError();
        };
        cudaError_t err;
        // Default: 3 producers + 7 consumers per CTA around 17 slots of 16 rows, 8 level-1 blocks
        // in flight per producer, two CTAs (20 warps) per SM.  Measured per 4 096 planted COCO
        // images / per 256 dense-random ones (MGD_DECODE_SHAPE selects the alternatives):
        //   <2,4,8,32>     1.24 / 0.50 ms      <3,6,7,32>      1.16 ms
        //   <3,6,15,16>    1.08 / 0.51 ms      <3,6,17,16,8>   1.08 / 0.50 ms
        //   <3,7,17,16,8>  1.03 / 0.47 ms      <3,8,17,16,8>   1.06 / 0.46 ms (80 registers)
        //   <4,6,14,16>    1.56 / 0.79 ms (a fourth producer starves the ring of slots)
        if (env_shape == 1) err = run(ws::Shape<3, 6, 17, 16, 8>());
        else if (env_shape == 3) err = run(ws::Shape<3, 8, 17, 16, 8>());
        else if (env_shape == 4) err = run(ws::Shape<3, 6, 7, 32>());
        else if (env_shape == 5) err = run(ws::Shape<2, 4, 8, 32>());
        else err = run(ws::Shape<3, 7, 17, 16, 8>());
        if (err != cudaErrorInvalidConfiguration) return err;
        cudaGetLastError();                 // too wide even for one CTA per SM: generic kernel
    }
    static int env_pool = -1;
    if (env_pool < 0) { const char* e = getenv("MGD_DECODE_POOL_ROWS"); env_pool = e ? atoi(e) : 0; }
    int pool_rows = env_pool == 32 || env_pool == 24 || env_pool == 16 ? env_pool : 32;
    // wide heads (hundreds of classes): smaller pools, down to 4 rows per warp (D <= ~1750)
    while (pool_rows > 16 && (size_t)kWarpsPerCta * pool_rows * stride * sizeof(float) > 200 * 1024) pool_rows -= 8;
    while (pool_rows > 4 && (size_t)kWarpsPerCta * pool_rows * stride * sizeof(float) > 200 * 1024) pool_rows /= 2;
    const size_t smem = (size_t)kWarpsPerCta * pool_rows * stride * sizeof(float);
    if (smem > 220 * 1024) return cudaErrorInvalidConfiguration;     // reported as MGD_ERR_UNSUPPORTED
    int ctas_per_sm = (int)((224 * 1024) / (smem + 2048));
    const int reg_limit = pool_rows >= 32 ? 2 : (pool_rows >= 24 ? 3 : (pool_rows >= 16 ? 4 : 2));
    if (ctas_per_sm > reg_limit) ctas_per_sm = reg_limit;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    long long blocks = 0;
    for (int l = 0; l < g.L; ++l) blocks += (a.rows_in_layer[l] + 31) / 32;
    long long grid = (long long)num_sms * ctas_per_sm;
    const long long needed = (blocks + kWarpsPerCta - 1) / kWarpsPerCta;
    if (grid > needed) grid = needed;
    if (grid < 1) grid = 1;
    cudaError_t err;
    prof_mark_begin(PROF_DECODE_COMPACT, stream);
    auto run = [&](auto kernel) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        kernel<<<(unsigned)grid, kThreads, smem, stream>>>(a, stride);
        return cudaSuccess;
    };
    if (fast) {
        if (pool_rows == 32) err = run(decode_compact_kernel<true, 32>);
        else if (pool_rows == 24) err = run(decode_compact_kernel<true, 24>);
        else if (pool_rows == 16) err = run(decode_compact_kernel<true, 16>);
        else if (pool_rows == 8) err = run(decode_compact_kernel<true, 8>);
        else err = run(decode_compact_kernel<true, 4>);
    } else {
        if (pool_rows == 32) err = run(decode_compact_kernel<false, 32>);
        else if (pool_rows == 24) err = run(decode_compact_kernel<false, 24>);
        else if (pool_rows == 16) err = run(decode_compact_kernel<false, 16>);
        else if (pool_rows == 8) err = run(decode_compact_kernel<false, 8>);
        else err = run(decode_compact_kernel<false, 4>);
    }
    if (err != cudaSuccess) return err;
    prof_mark_end(PROF_DECODE_COMPACT, stream);
    return cudaGetLastError();
}

cudaError_t launch_decode_dense(const DecodeArgs& a, const int* image_hw, double* out,
                                cudaStream_t stream)
{
    const HeadGeom& g = a.g;
    int dmax = 0;
    for (int l = 0; l < g.L; ++l) dmax = g.D[l] > dmax ? g.D[l] : dmax;
    const int row_floats = (dmax + 3) & ~3;
    const size_t smem = (size_t)(kDenseThreads / 8) * row_floats * 4;
    cudaError_t err = cudaFuncSetAttribute(decode_dense_kernel,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    const long long rows = (long long)a.B * g.cells;
    long long grid = (rows + kDenseThreads / 8 - 1) / (kDenseThreads / 8);
    if (grid > 148 * 8) grid = 148 * 8;
    if (grid < 1) grid = 1;
    prof_mark_begin(PROF_OTHER, stream);
    decode_dense_kernel<<<(unsigned)grid, kDenseThreads, smem, stream>>>(a, image_hw, out, row_floats);
    prof_mark_end(PROF_OTHER, stream);
    return cudaGetLastError();
}
