// Dense head decode + score threshold + candidate compaction for sm_100a.
//
// Replaces MultiGridDecoder.decode_predictions / correct_boxes and the threshold
// step of handle_predictions (reference multigriddet/postprocess/
// multigrid_decode.py:100-183, 185-235, 262-278) without ever materialising the
// reference's (B, 7581, 85) float64 tensor.
//
// decode_compact_kernel -- persistent CTAs, HBM-bound (reads cells*D*4 bytes):
//   * the head tensor streams through shared memory in tiles of whole cell rows,
//     moved by the TMA bulk-copy engine (cp.async.bulk + mbarrier, multi-stage
//     ring), so no thread spends registers or issue slots on loads;
//   * phase 1: one compare per row on the raw objectness logit.  score <=
//     sigmoid(obj) because both softmax maxima are <= 1, so a row whose logit is
//     below logit(confidence) minus a margin cannot become a candidate;
//   * phase 2: the survivors (about 10% of rows on a trained head) are evaluated
//     exactly, eight lanes per row, in the reference's float32 operation order:
//     softmax as exp(x - max) / sum with NumPy's pairwise 8-accumulator summation
//     order (the eight lanes ARE the eight accumulators), glibc-equivalent expf
//     (libm_emul.h), first-maximum argmax on the probabilities, score =
//     (obj * anchor) * class, threshold in float64 like `score >= confidence`;
//   * candidates get their box in float64 ((xy + cell) / grid, anchor * exp(wh),
//     letterbox correction with float32 constants) and are appended to the
//     per-image candidate list with one atomic per candidate.
#include <math.h>
#include "common.cuh"
#include "libm_emul.h"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxStages = 8;

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
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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

// ---- NumPy float32 add.reduce order, one octet of lanes = the 8 accumulators --
// x: shared-memory array of n floats; j = lane within the octet (0..7).
__device__ float np_sum_octet(const float* x, int n, int j)
{
    if (n < 8) {
        float res = 0.f;
        for (int i = 0; i < n; ++i) res = __fadd_rn(res, x[i]);
        return res;
    }
    if (n <= 128) {
        float r = x[j];
        const int body = n - (n & 7);
        for (int i = 8 + j; i < body; i += 8) r = __fadd_rn(r, x[i]);
        r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));   // (r0+r1) (r2+r3) ...
        r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));   // ((r0+r1)+(r2+r3)) ...
        r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
        for (int i = body; i < n; ++i) r = __fadd_rn(r, x[i]);
        return r;
    }
    int n2 = n / 2;
    n2 -= n2 & 7;
    const float lo = np_sum_octet(x, n2, j);
    const float hi = np_sum_octet(x + n2, n - n2, j);
    return __fadd_rn(lo, hi);
}

// Max probability and its first index over x[0..n): softmax (scipy: exp(x-max)/sum)
// or element-wise expit.  Overwrites x with the exponentials / probabilities.
// All 32 lanes of the warp must call this together (full-mask shuffles).
__device__ void octet_probs(float* x, int n, int j, bool use_softmax, const uint64_t* tab,
                            float& pmax, int& arg)
{
    if (use_softmax) {
        float m = -INFINITY;
        for (int i = j; i < n; i += 8) m = fmaxf(m, x[i]);
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
        for (int i = j; i < n; i += 8) x[i] = mgd_expf_tab(__fsub_rn(x[i], m), tab);
        __syncwarp();
        const float s = np_sum_octet(x, n, j);
        pmax = __fdiv_rn(1.0f, s);              // the maximum's exponential is exactly 1
        int first = INT_MAX;
        for (int i = j; i < n; i += 8) {
            const float e = x[i];
            if (e >= 0.99999f && __fdiv_rn(e, s) == pmax) { first = i; break; }
        }
        first = min(first, __shfl_xor_sync(0xffffffffu, first, 1));
        first = min(first, __shfl_xor_sync(0xffffffffu, first, 2));
        first = min(first, __shfl_xor_sync(0xffffffffu, first, 4));
        arg = first;
    } else {
        float best = -INFINITY;
        int first = INT_MAX;
        for (int i = j; i < n; i += 8) {
            const float q = mgd_expitf_tab(x[i], tab);
            x[i] = q;
            if (q > best) { best = q; first = i; }
        }
        #pragma unroll
        for (int d = 1; d <= 4; d <<= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, d);
            const int oi = __shfl_xor_sync(0xffffffffu, first, d);
            if (ob > best || (ob == best && oi < first)) { best = ob; first = oi; }
        }
        pmax = best;
        arg = first;
    }
}

struct Letterbox { float off_w, off_h, sc_w, sc_h, img_w, img_h; };

// multigrid_decode.py:205-216 in float32
__device__ __forceinline__ Letterbox letterbox_consts(int in_h, int in_w, int ih_i, int iw_i)
{
    const float mh = (float)in_h, mw = (float)in_w, ih = (float)ih_i, iw = (float)iw_i;
    const float ratio = fminf(__fdiv_rn(mh, ih), __fdiv_rn(mw, iw));
    const float nh = rintf(__fmul_rn(ih, ratio)), nw = rintf(__fmul_rn(iw, ratio));
    Letterbox lb;
    lb.off_h = __fdiv_rn(__fdiv_rn(__fsub_rn(mh, nh), 2.0f), mh);
    lb.off_w = __fdiv_rn(__fdiv_rn(__fsub_rn(mw, nw), 2.0f), mw);
    lb.sc_h = __fdiv_rn(mh, nh);
    lb.sc_w = __fdiv_rn(mw, nw);
    lb.img_w = iw;
    lb.img_h = ih;
    return lb;
}

// Box of one cell in float64: multigrid_decode.py:151-163 then :219-228.
// x: the raw row (channels 0..3 untouched), ga: global anchor index.
__device__ __forceinline__ void decode_box(const HeadGeom& g, const float* x, int layer, int ga,
                                           int r, int c, const Letterbox* lb,
                                           const uint64_t* tab, double out[4])
{
    const float ux = __fmul_rn(0.15f, x[0]), uy = __fmul_rn(0.15f, x[1]);
    // np.tanh float32: correctly rounded here (libm tanhf is within 2 ulp of this)
    const float ax = __fadd_rn((float)tanh((double)ux), mgd_expitf_tab(ux, tab));
    const float ay = __fadd_rn((float)tanh((double)uy), mgd_expitf_tab(uy, tab));
    double bx = __ddiv_rn(__dadd_rn((double)ax, (double)c), (double)g.gh[layer]);   // :154-155
    double by = __ddiv_rn(__dadd_rn((double)ay, (double)r), (double)g.gw[layer]);
    double bw, bh;
    const float ew = mgd_expf_tab(x[2], tab), eh = mgd_expf_tab(x[3], tab);
    if (!g.anchors_f64) {
        const float w32 = __fmul_rn(g.anc32[ga][0], ew), h32 = __fmul_rn(g.anc32[ga][1], eh);
        bw = (double)(float)__ddiv_rn((double)w32, (double)g.in_h);    // :163 in-place on f32
        bh = (double)(float)__ddiv_rn((double)h32, (double)g.in_w);
    } else {
        bw = __ddiv_rn(__dmul_rn(g.anc64[ga][0], (double)ew), (double)g.in_h);
        bh = __ddiv_rn(__dmul_rn(g.anc64[ga][1], (double)eh), (double)g.in_w);
    }
    if (lb) {
        bx = __dmul_rn(__dsub_rn(bx, (double)lb->off_w), (double)lb->sc_w);        // :219
        by = __dmul_rn(__dsub_rn(by, (double)lb->off_h), (double)lb->sc_h);
        bw = __dmul_rn(bw, (double)lb->sc_w);                                      // :220
        bh = __dmul_rn(bh, (double)lb->sc_h);
        bx = __dsub_rn(bx, __ddiv_rn(bw, 2.0));                                    // :223
        by = __dsub_rn(by, __ddiv_rn(bh, 2.0));
        bx = __dmul_rn(bx, (double)lb->img_w);                                     // :227-228
        by = __dmul_rn(by, (double)lb->img_h);
        bw = __dmul_rn(bw, (double)lb->img_w);
        bh = __dmul_rn(bh, (double)lb->img_h);
    }
    out[0] = bx; out[1] = by; out[2] = bw; out[3] = bh;
}

// Dynamic shared memory layout: [stages][rows_per_tile * Dmax] floats, then the
// survivor list.  kUseTma=false copies tiles with ordinary loads (any D / alignment).
template <bool kUseTma>
__global__ void __launch_bounds__(kThreads)
decode_compact_kernel(const __grid_constant__ DecodeArgs a, int n_stages, int stage_floats)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full_bar[kMaxStages];
    __shared__ uint64_t s_tab[MGD_EXP2F_N];
    __shared__ int s_count[2];

    const HeadGeom& g = a.g;
    float* stage0 = reinterpret_cast<float*>(smem_raw);
    int* s_list = reinterpret_cast<int*>(smem_raw + (size_t)n_stages * stage_floats * sizeof(float));
    float* s_dummy = reinterpret_cast<float*>(s_list + a.rows_per_tile);

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int j = tid & 7;                    // lane within the octet
    const int oct = lane >> 3;                // octet within the warp
    const int warp = tid >> 5;
    const long long total_tiles = a.tile_first[g.L];

    if (tid < MGD_EXP2F_N) s_tab[tid] = mgd_exp2f_tab[tid];
    for (int i = tid; i < stage_floats / a.rows_per_tile; i += kThreads) s_dummy[i] = 0.f;
    if (tid == 0) {
        s_count[0] = 0; s_count[1] = 0;
        if (kUseTma) {
            for (int s = 0; s < n_stages; ++s) mbar_init(&full_bar[s], 1);
            fence_mbar_init();
        }
    }
    __syncthreads();

    auto tile_geom = [&](long long tile, int& layer, long long& row0, int& rows) {
        int l = 0;
        while (l + 1 < g.L && tile >= a.tile_first[l + 1]) ++l;
        layer = l;
        row0 = (tile - a.tile_first[l]) * a.rows_per_tile;
        const long long left = a.rows_in_layer[l] - row0;
        rows = (int)(left < a.rows_per_tile ? left : a.rows_per_tile);
    };
    auto issue = [&](long long seq) {          // thread 0 only
        const long long tile = blockIdx.x + seq * gridDim.x;
        if (tile >= total_tiles) return;
        int layer, rows; long long row0;
        tile_geom(tile, layer, row0, rows);
        const int s = (int)(seq % n_stages);
        const uint32_t bytes = (uint32_t)rows * g.D[layer] * sizeof(float);
        mbar_arrive_expect_tx(&full_bar[s], bytes);
        bulk_g2s(stage0 + (size_t)s * stage_floats, a.pred[layer] + row0 * g.D[layer], bytes,
                 &full_bar[s]);
    };
    if (kUseTma && tid == 0)
        for (int s = 0; s < n_stages; ++s) issue(s);

    for (long long seq = 0;; ++seq) {
        const long long tile = blockIdx.x + seq * gridDim.x;
        if (tile >= total_tiles) break;
        int layer, rows; long long row0;
        tile_geom(tile, layer, row0, rows);
        const int D = g.D[layer];
        const int A = g.na[layer];
        const int s = (int)(seq % n_stages);
        float* buf = stage0 + (size_t)s * stage_floats;
        int* cnt = &s_count[seq & 1];

        if (kUseTma) {
            mbar_wait(&full_bar[s], (uint32_t)((seq / n_stages) & 1));
        } else {
            const float* src = a.pred[layer] + row0 * D;
            for (int i = tid; i < rows * D; i += kThreads) buf[i] = __ldg(src + i);
            __syncthreads();
        }

        // ---- phase 1: objectness prefilter, one compare per row ----------------
        for (int r = tid; r < rows; r += kThreads)
            if (buf[r * D + 4] >= a.obj_logit_min) s_list[atomicAdd(cnt, 1)] = r;
        __syncthreads();
        const int n_surv = *cnt;
        if (tid == 0) s_count[(seq + 1) & 1] = 0;

        // ---- phase 2: exact evaluation, one octet of lanes per surviving row ---
        for (int base = warp * 4; base < n_surv; base += (kThreads / 32) * 4) {
            const int slot = base + oct;
            const bool live = slot < n_surv;
            const int r = live ? s_list[slot] : 0;
            // Octets without a row still run the (full-mask) shuffles below; they
            // work on a scratch row so they never touch a row another octet rewrites.
            float* x = live ? buf + r * D : s_dummy;
            float pa, pc; int ka, kc;
            octet_probs(x + 5, A, j, a.use_softmax != 0, s_tab, pa, ka);
            octet_probs(x + 5 + A, g.C, j, a.use_softmax != 0, s_tab, pc, kc);
            float score = mgd_expitf_tab(x[4], s_tab);                       // :147
            if (a.rescore) score = __fmul_rn(__fmul_rn(score, pa), pc);      // :170
            if (live && j == 0 && (double)score >= a.confidence) {           // :271
                const long long grow = row0 + r;
                const int cells_l = g.gh[layer] * g.gw[layer];
                const int b = (int)(grow / cells_l);
                const int cell = (int)(grow - (long long)b * cells_l);
                const int rr = cell / g.gw[layer], cc = cell - rr * g.gw[layer];
                const int ih = a.image_hw ? a.image_hw[2 * b] : g.in_h;
                const int iw = a.image_hw ? a.image_hw[2 * b + 1] : g.in_w;
                const Letterbox lb = letterbox_consts(g.in_h, g.in_w, ih, iw);
                double box[4];
                decode_box(g, x, layer, g.anchor_first[layer] + ka, rr, cc, &lb, s_tab, box);
                Cand cd;
                cd.x = box[0]; cd.y = box[1]; cd.w = box[2]; cd.h = box[3];
                cd.score = (double)score;
                cd.index = g.cell_off[layer] + cell;
                cd.cls = kc;
                const int pos = atomicAdd(a.counts + b, 1);
                a.cand[(size_t)b * g.cells + pos] = cd;
            }
        }
        __syncthreads();                       // stage s and the list are free again
        if (kUseTma && tid == 0) issue(seq + n_stages);
    }
}

// ---- dense decode (decode_predictions API), one octet per row, not a hot path --
__global__ void __launch_bounds__(kThreads)
decode_dense_kernel(const __grid_constant__ DecodeArgs a, const int* image_hw, double* out,
                    int row_floats)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint64_t s_tab[MGD_EXP2F_N];
    const HeadGeom& g = a.g;
    const int tid = threadIdx.x, j = tid & 7;
    float* x = reinterpret_cast<float*>(smem_raw) + (size_t)(tid >> 3) * row_floats;
    if (tid < MGD_EXP2F_N) s_tab[tid] = mgd_exp2f_tab[tid];
    __syncthreads();
    const long long total_rows = (long long)a.B * g.cells;
    const long long octs = (long long)gridDim.x * (kThreads / 8);
    // uniform trip count per warp: every lane iterates while ANY octet of the grid might
    const long long iters = (total_rows + octs - 1) / octs;
    for (long long it = 0; it < iters; ++it) {
        long long q = it * octs + (long long)blockIdx.x * (kThreads / 8) + (tid >> 3);
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
        __syncwarp();
        float pa, pc; int ka, kc;
        octet_probs(x + 5, A, j, a.use_softmax != 0, s_tab, pa, ka);
        octet_probs(x + 5 + A, g.C, j, a.use_softmax != 0, s_tab, pc, kc);
        __syncwarp();
        float score = mgd_expitf_tab(x[4], s_tab);
        if (a.rescore) score = __fmul_rn(__fmul_rn(score, pa), pc);
        const int rr = cell / g.gw[layer], cc = cell - rr * g.gw[layer];
        double box[4];
        Letterbox lb;
        if (image_hw) lb = letterbox_consts(g.in_h, g.in_w, image_hw[2 * b], image_hw[2 * b + 1]);
        decode_box(g, x, layer, g.anchor_first[layer] + ka, rr, cc, image_hw ? &lb : nullptr,
                   s_tab, box);
        if (live) {
            double* o = out + (size_t)q * (5 + g.C);
            if (j < 4) o[j] = box[j];
            if (j == 4) o[4] = (double)score;
            float s = 1.0f;
            if (a.use_softmax) s = np_sum_octet(x + 5 + A, g.C, j);
            for (int i = j; i < g.C; i += 8) {
                const float e = x[5 + A + i];
                o[5 + i] = (double)(a.use_softmax ? __fdiv_rn(e, s) : e);
            }
        } else if (a.use_softmax) {
            (void)np_sum_octet(x + 5 + A, g.C, j);
        }
        __syncwarp();
    }
}

}  // namespace

cudaError_t launch_decode(const DecodeArgs& a_in, int num_sms, cudaStream_t stream)
{
    DecodeArgs a = a_in;
    const HeadGeom& g = a.g;
    int dmax = 0;
    bool tma_ok = true;
    for (int l = 0; l < g.L; ++l) {
        dmax = g.D[l] > dmax ? g.D[l] : dmax;
        tma_ok = tma_ok && (g.D[l] % 4 == 0) &&
                 ((reinterpret_cast<uintptr_t>(a.pred[l]) & 15) == 0);
    }
    // tile = whole rows; a stage holds up to rows_per_tile rows of the widest layer
    static int env_rows = -1, env_stages = -1, env_ctas = -1;
    if (env_rows < 0) {
        const char* e;
        env_rows = (e = getenv("MGD_DECODE_TILE_ROWS")) ? atoi(e) : 0;
        env_stages = (e = getenv("MGD_DECODE_STAGES")) ? atoi(e) : 0;
        env_ctas = (e = getenv("MGD_DECODE_CTAS_PER_SM")) ? atoi(e) : 0;
    }
    int ctas_per_sm = env_ctas > 0 ? env_ctas : 2;
    int n_stages = env_stages > 0 ? env_stages : 3;
    if (n_stages > kMaxStages) n_stages = kMaxStages;
    const size_t budget = (size_t)(220 * 1024) / ctas_per_sm - 2048;
    int rows = env_rows > 0 ? env_rows : 64;
    while (rows > 8 && (size_t)n_stages * rows * dmax * 4 + (size_t)(rows + dmax + 32) * 4 > budget) rows /= 2;
    if ((size_t)n_stages * rows * dmax * 4 + (size_t)(rows + dmax + 32) * 4 > 220 * 1024) return cudaErrorInvalidValue;
    if (!tma_ok) n_stages = 1;
    a.rows_per_tile = rows;
    long long tiles = 0;
    for (int l = 0; l < g.L; ++l) {
        a.rows_in_layer[l] = (long long)a.B * g.gh[l] * g.gw[l];
        a.tile_first[l] = tiles;
        tiles += (a.rows_in_layer[l] + rows - 1) / rows;
    }
    a.tile_first[g.L] = tiles;
    // rows is a power of two >= 8 here, so every stage stays 32-byte aligned;
    // stage_floats / rows == dmax is the scratch-row length the kernel derives
    const int stage_floats = rows * dmax;
    const size_t smem = (size_t)n_stages * stage_floats * 4 + (size_t)(rows + dmax + 32) * 4;
    long long grid = (long long)num_sms * ctas_per_sm;
    if (grid > tiles) grid = tiles;
    if (grid < 1) grid = 1;
    cudaError_t err;
    prof_mark_begin(PROF_DECODE_COMPACT, stream);
    if (tma_ok) {
        err = cudaFuncSetAttribute(decode_compact_kernel<true>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        decode_compact_kernel<true><<<(unsigned)grid, kThreads, smem, stream>>>(a, n_stages, stage_floats);
    } else {
        err = cudaFuncSetAttribute(decode_compact_kernel<false>,
                                   cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (err != cudaSuccess) return err;
        decode_compact_kernel<false><<<(unsigned)grid, kThreads, smem, stream>>>(a, n_stages, stage_floats);
    }
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
    const size_t smem = (size_t)(kThreads / 8) * row_floats * 4;
    cudaError_t err = cudaFuncSetAttribute(decode_dense_kernel,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    const long long rows = (long long)a.B * g.cells;
    long long grid = (rows + kThreads / 8 - 1) / (kThreads / 8);
    if (grid > 148 * 8) grid = 148 * 8;
    if (grid < 1) grid = 1;
    prof_mark_begin(PROF_OTHER, stream);
    decode_dense_kernel<<<(unsigned)grid, kThreads, smem, stream>>>(a, image_hw, out, row_floats);
    prof_mark_end(PROF_OTHER, stream);
    return cudaGetLastError();
}
