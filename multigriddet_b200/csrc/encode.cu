// Multi-grid y_true target encoder for sm_100a.
//
// Replaces the O(B*N*9) Python loop of preprocess_true_boxes
// (reference multigriddet/data/generators.py:3393-3473) with two kernels:
//
//   encode_assign_kernel  one CTA per image, thread per ground-truth box.
//       Anchor matching (generators.py:2486-2494, 2529-2532) and the 3x3 scatter
//       with the reference's sequential overwrite rule (:3454-3472), resolved
//       without any ordering between boxes:
//         cover[cell] = min box index whose 3x3 block contains the cell
//                       (== "cell occupied before box t"  <=>  cover[cell] < t,
//                        because the first box covering a cell always writes it)
//         a box skips a candidate iff cover[cell] < t and it has written >= 3
//         owner[cell] = max box index among the boxes that write the cell
//       Both tables live in shared memory (atomicMin / atomicMax); the result is
//       a per-cell owner code (box record, neighbour id) written once, coalesced.
//
//   encode_fill_kernel    grid-wide streaming writer.  A warp owns tiles of whole
//       rows whose size is a multiple of 512 B: it stores the tile as zeros with 32
//       consecutive float4 per instruction (four full 128-byte lines), then patches
//       the seven non-zero channels of the rows that have an owner while the lines
//       are still in L2 -- DRAM sees every byte exactly once, there is no separate
//       memset pass over the tensor.  The per-row owner codes of a tile arrive in
//       one coalesced load, prefetched a tile ahead.  HBM-bound: algorithmic bytes =
//       cells*D*4 written + 20 B/box read.
//
// All arithmetic that decides an integer (anchor, layer, cell, skip) uses IEEE
// single/double operations in the reference's order; the file is compiled with
// -fmad=false so nothing is contracted.
#include <limits.h>
#include <math.h>
#include <string.h>
#include "common.cuh"

namespace {

constexpr int kAssignThreads = 256;
constexpr int kFillThreads = 256;

// The owner table travels from the assign kernel to the writer through L2 (32 MB per chunk).
// Stored with the L2 evict_last priority: the writer's 3 GB store stream otherwise pushes it
// out to DRAM before it is read back (ncu: 38 MB of DRAM reads per launch), and those reads
// cost the write stream 4 % of its rate (0.418 -> 0.399 ms per 1 024 images; the lookup as a
// whole costs 11 %: 0.372 ms without it, MGD_FILL_DEBUG=1).  A matching evict_first hint on
// the read changes nothing.
__device__ __forceinline__ void st_table_keep(int* p, int v)
{
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("st.global.L2::cache_hint.b32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}

// generators.py:2486-2494 + np.round(iol, 3) + first maximum (lowest global index)
__device__ __forceinline__ int match_anchor(const HeadGeom& g, float bw, float bh)
{
    int best = 0;
    if (!g.anchors_f64) {
        float top = -INFINITY;
        const float box_area = __fmul_rn(bw, bh);
        for (int i = 0; i < g.K; ++i) {
            const float aw = g.anc32[i][0], ah = g.anc32[i][1];
            const float inter = __fmul_rn(fminf(bw, aw), fminf(bh, ah));
            const float iol = __fdiv_rn(inter, fmaxf(box_area, __fmul_rn(aw, ah)));
            // round(x,3) = rint(x*1000)/1000; the division is monotone and injective
            // on these integers, so the argmax can be taken on rint(x*1000)
            const float r = rintf(__fmul_rn(iol, 1000.0f));
            if (r > top) { top = r; best = i; }
        }
    } else {
        double top = -INFINITY;
        const double box_area = (double)__fmul_rn(bw, bh);   // float32 product, :2489
        for (int i = 0; i < g.K; ++i) {
            const double aw = g.anc64[i][0], ah = g.anc64[i][1];
            const double inter = __dmul_rn(fmin((double)bw, aw), fmin((double)bh, ah));
            const double iol = __ddiv_rn(inter, fmax(box_area, __dmul_rn(aw, ah)));
            const double r = rint(__dmul_rn(iol, 1000.0));
            if (r > top) { top = r; best = i; }
        }
    }
    return best;
}

// tf_preprocess_true_boxes (generators.py:2811-2931): float32 IoL with the Keras epsilon
// in the denominator, no rounding, first maximum over the layer-major anchor list
__device__ __forceinline__ int match_anchor_tf(const HeadGeom& g, float bw, float bh)
{
    int best = 0;
    float top = -INFINITY;
    const float box_area = __fmul_rn(bw, bh);
    for (int i = 0; i < g.K; ++i) {
        const float aw = g.anc32[i][0], ah = g.anc32[i][1];
        const float inter = __fmul_rn(fminf(bw, aw), fminf(bh, ah));
        const float iol = __fdiv_rn(inter, __fadd_rn(fmaxf(box_area, __fmul_rn(aw, ah)), 1e-7f));
        if (iol > top) { top = iol; best = i; }
    }
    return best;
}

// packed per-box placement kept in shared memory between the phases
__device__ __forceinline__ int pack_place(int layer, int col, int row)
{
    // anything further than one cell outside the grid has no in-bounds candidate
    col = min(max(col, -2), 16380) + 2;
    row = min(max(row, -2), 16380) + 2;
    return (layer << 28) | (row << 14) | col;
}

__global__ void __launch_bounds__(kAssignThreads)
encode_assign_kernel(const __grid_constant__ EncodeArgs a)
{
    extern __shared__ int sm[];
    // (programmatic dependent launch, see encode_fill_kernel: the assign kernels of the chunks of
    //  one mgd_encode_decode_nms call are independent of each other)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const HeadGeom& g = a.g;
    // the two per-cell tables live in shared memory; heads too large for that (e.g. a
    // stride-2 layer) fall back to a per-image slice of global scratch (L2-resident)
    const bool big = a.big_tables != nullptr;
    int* cover = big ? a.big_tables + (size_t)blockIdx.x * 2 * g.cells : sm;   // [cells]
    int* owner = cover + g.cells;                                            // [cells]
    int* place = big ? sm : owner + g.cells;                                 // [N]
    __shared__ unsigned int s_stat[3];
    __shared__ int s_status;

    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    for (int i = tid; i < g.cells; i += kAssignThreads) { cover[i] = INT_MAX; owner[i] = -1; }
    if (tid < 3) s_stat[tid] = 0;
    if (tid == 0) s_status = 0;
    __syncthreads();

    unsigned int n_valid = 0, n_skipped = 0, n_pos = 0;
    int status = 0;

    if (a.tf_compat) {
        // ---- TensorFlow semantics (generators.py:2696-3390), one pass ----------------------
        // exact centre, unrounded float32 IoL, every in-bounds cell of the 3x3 block written
        // (the occupancy test reads a still-empty tensor, :3245), duplicates resolved like
        // tensor_scatter_nd_update on the CPU: the highest box index wins.  Stored xy =
        // [-kj + frac(cy), -ki + frac(cx)] (:3337-3339): rec.fx / rec.fy hold the swapped
        // fractions so the fill kernel is the same.  No class-range error: one_hot leaves
        // the class channels zero for ids outside [0, C) (:3351) -> hot_class = objectness.
        for (int t = tid; t < a.N; t += kAssignThreads) {
            const float* bx = a.boxes + ((size_t)b * a.N + t) * 5;
            const float x1 = bx[0], y1 = bx[1], x2 = bx[2], y2 = bx[3], cf = bx[4];
            const float bw = __fsub_rn(x2, x1), bh = __fsub_rn(y2, y1);          // :2731
            if (!(__fmul_rn(bw, bh) > 0.0f)) continue;                            // :2757
            ++n_valid;
            const float cxp = __fmul_rn(__fadd_rn(x1, x2), 0.5f);                 // :2730 (/2.0 exact)
            const float cyp = __fmul_rn(__fadd_rn(y1, y2), 0.5f);
            const int ga = match_anchor_tf(g, bw, bh);
            int layer = 0;
            while (layer + 1 < g.L && ga >= g.anchor_first[layer + 1]) ++layer;
            const int k = ga - g.anchor_first[layer];
            const int gh = g.gh[layer], gw = g.gw[layer];
            const float cx = __fmul_rn(cxp, __fdiv_rn((float)gw, (float)g.in_w)); // :2960-2965
            const float cy = __fmul_rn(cyp, __fdiv_rn((float)gh, (float)g.in_h));
            const int col = (int)fminf(fmaxf(cx, -4.0f), 1.0e6f);                 // tf.cast: truncation
            const int row = (int)fminf(fmaxf(cy, -4.0f), 1.0e6f);
            BoxRec rec;
            rec.fx = (double)__fsub_rn(cy, (float)row);                           // channel 0 gets ty
            rec.fy = (double)__fsub_rn(cx, (float)col);                           // channel 1 gets tx
            const float rw = fmaxf(__fdiv_rn(bw, g.anc32[ga][0]), 1e-3f);         // :3330-3333
            const float rh = fmaxf(__fdiv_rn(bh, g.anc32[ga][1]), 1e-3f);
            rec.tw = (float)log((double)rw);
            rec.th = (float)log((double)rh);
            rec.hot_anchor = 5 + k;
            // tf.cast(float -> int32) truncates; out-of-range ids light no class channel
            const int cls = (cf >= 0.0f && cf < 2.0e9f) ? (int)cf : -1;
            rec.hot_class = (cls >= 0 && cls < g.C) ? 5 + g.na[layer] + cls : 4;
            a.recs[(size_t)b * a.N + t] = rec;
            #pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                #pragma unroll
                for (int dy = -1; dy <= 1; ++dy) {
                    const int cc = col + dx, rr = row + dy;
                    if (cc >= 0 && cc < gw && rr >= 0 && rr < gh)
                        atomicMax(&owner[g.cell_off[layer] + rr * gw + cc],
                                  t * 16 + (dx + 1) * 3 + (dy + 1));
                }
            }
        }
        __syncthreads();
    } else {
        // ---- phase A: per-box record + cover table --------------------------------
        for (int t = tid; t < a.N; t += kAssignThreads) {
            const float* bx = a.boxes + ((size_t)b * a.N + t) * 5;
            const float x1 = bx[0], y1 = bx[1], x2 = bx[2], y2 = bx[3], cf = bx[4];
            if (!(cf < (float)g.C)) status |= 1;                         // :3409, every row
            const float bw = __fsub_rn(x2, x1), bh = __fsub_rn(y2, y1);  // :3416
            int pl = -1;
            if (!(__fmul_rn(bw, bh) <= 0.0f)) {                          // :3431
                ++n_valid;
                const int cls = (int)cf;                                 // astype('int32')
                if (cls < 0) status |= 2;
                // (x1+x2)//2 on float32 == floor((x1+x2)/2): the halving is exact
                const float cxp = floorf(__fmul_rn(__fadd_rn(x1, x2), 0.5f));   // :3415
                const float cyp = floorf(__fmul_rn(__fadd_rn(y1, y2), 0.5f));
                const int ga = match_anchor(g, bw, bh);
                int layer = 0;
                while (layer + 1 < g.L && ga >= g.anchor_first[layer + 1]) ++layer;
                const int k = ga - g.anchor_first[layer];
                // :3438-3439  f32 * (G/S as float64) -> float64 (NumPy 2 promotion)
                const double gx = __dmul_rn((double)cxp, __ddiv_rn((double)g.gh[layer], (double)g.in_h));
                const double gy = __dmul_rn((double)cyp, __ddiv_rn((double)g.gw[layer], (double)g.in_w));
                const int col = (int)fmin(fmax(gx, -4.0), 1.0e6);        // int(): truncation
                const int row = (int)fmin(fmax(gy, -4.0), 1.0e6);
                BoxRec rec;
                rec.fx = __dsub_rn(gx, (double)col);
                rec.fy = __dsub_rn(gy, (double)row);
                if (!g.anchors_f64) {                                    // :3446-3449
                    const float rw = __fdiv_rn(bw, g.anc32[ga][0]);
                    const float rh = __fdiv_rn(bh, g.anc32[ga][1]);
                    // NumPy's float32 log is libm logf there; log in double rounded once
                    // to float reproduces it (<= 1 ulp away in ~0.4% of inputs)
                    rec.tw = (float)log(rw < 1e-3f ? 1e-3 : (double)rw);
                    rec.th = (float)log(rh < 1e-3f ? 1e-3 : (double)rh);
                } else {
                    const double rw = __ddiv_rn((double)bw, g.anc64[ga][0]);
                    const double rh = __ddiv_rn((double)bh, g.anc64[ga][1]);
                    rec.tw = (float)log(rw < 1e-3 ? 1e-3 : rw);
                    rec.th = (float)log(rh < 1e-3 ? 1e-3 : rh);
                }
                rec.hot_anchor = 5 + k;
                rec.hot_class = 5 + g.na[layer] + max(cls, 0);
                a.recs[(size_t)b * a.N + t] = rec;
                pl = pack_place(layer, col, row);
                const int g0 = g.gh[layer], g1 = g.gw[layer];
                const int c0 = (pl & 0x3fff) - 2, r0 = ((pl >> 14) & 0x3fff) - 2;
                #pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    #pragma unroll
                    for (int dy = -1; dy <= 1; ++dy) {
                        const int cc = c0 + dx, rr = r0 + dy;
                        if (cc >= 0 && cc < g0 && rr >= 0 && rr < g1)
                            atomicMin(&cover[g.cell_off[layer] + rr * g1 + cc], t);
                    }
                }
            }
            place[t] = pl;
        }
        __syncthreads();

        // ---- phase B: sequential skip rule per box, owner = last writer -----------
        for (int t = tid; t < a.N; t += kAssignThreads) {
            const int pl = place[t];
            if (pl < 0) continue;
            const int layer = pl >> 28;
            const int c0 = (pl & 0x3fff) - 2, r0 = ((pl >> 14) & 0x3fff) - 2;
            const int g0 = g.gh[layer], g1 = g.gw[layer];
            int written = 0;
            #pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {                           // :3454 x outer
                #pragma unroll
                for (int dy = -1; dy <= 1; ++dy) {                       // :3456 y inner
                    const int cc = c0 + dx, rr = r0 + dy;
                    if (cc < 0 || cc >= g0 || rr < 0 || rr >= g1) continue;   // :3459-3462
                    const int cell = g.cell_off[layer] + rr * g1 + cc;
                    if (cover[cell] < t && written >= 3) { ++n_skipped; continue; }   // :3463
                    ++written;
                    atomicMax(&owner[cell], t * 16 + (dx + 1) * 3 + (dy + 1));
                }
            }
        }
        __syncthreads();
    }

    // ---- phase C: publish the owner codes, layer-major ------------------------
    const int rec_base = b * a.N * 16;
    for (int l = 0; l < g.L; ++l) {
        const int n = g.gh[l] * g.gw[l];
        int* dst = a.table + (size_t)a.B * g.cell_off[l] + (size_t)b * n;
        const int* src = owner + g.cell_off[l];
        for (int i = tid; i < n; i += kAssignThreads) {
            const int v = src[i];
            st_table_keep(dst + i, v < 0 ? -1 : v + rec_base);
            n_pos += (v >= 0);
        }
    }
    if (n_valid) atomicAdd(&s_stat[0], n_valid);
    if (n_skipped) atomicAdd(&s_stat[1], n_skipped);
    if (n_pos) atomicAdd(&s_stat[2], n_pos);
    if (status) atomicOr(&s_status, status);
    __syncthreads();
    if (tid < 3 && a.stats && s_stat[tid]) atomicAdd(&a.stats[tid], (unsigned long long)s_stat[tid]);
    if (tid == 0 && s_status) atomicOr(a.status, s_status);
}

// Per-layer plan of the streaming writer.  A warp tile is R whole rows with
// R * dv a multiple of 32 vector units, so every warp store covers 32 consecutive
// units (512 B with float4: four full 128-byte lines) and tiles never share a line.
struct FillPlan {
    int dv[MGD_MAX_LAYERS];                 // vector units per row = D / VEC
    int R[MGD_MAX_LAYERS];                  // rows per warp tile = 32 / gcd(dv, 32)
    long long tile_first[MGD_MAX_LAYERS + 1];
    long long rows[MGD_MAX_LAYERS];
    int debug;                              // MGD_FILL_DEBUG=1 (measurements): no owner lookup, no patches
};

template <int VEC> struct VecT;
template <> struct VecT<4> { using type = float4; };
template <> struct VecT<1> { using type = float; };

// n-th (0-based, n < 4) set bit of m, or 32 if m has fewer
__device__ __forceinline__ unsigned nth_bit(unsigned m, int n)
{
    #pragma unroll
    for (int q = 0; q < 3; ++q)
        if (q < n) m &= m - 1;
    return m ? (unsigned)(__ffs((int)m) - 1) : 32u;
}

template <int VEC>
__global__ void __launch_bounds__(kFillThreads)
encode_fill_kernel(const __grid_constant__ EncodeArgs a, const __grid_constant__ FillPlan p)
{
    using V = typename VecT<VEC>::type;
    // Programmatic dependent launch: the writer of the NEXT chunk does not depend on this one, so
    // when it was launched with the programmatic-serialization attribute (encode_fill_all) its
    // CTAs may start filling the SMs while this grid drains its last wave.  (No effect on
    // launches without the attribute.)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const HeadGeom& g = a.g;
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * kFillThreads + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * kFillThreads) >> 5;
    const long long total = p.tile_first[g.L];
    V zero;
    if constexpr (VEC == 4) zero = make_float4(0.f, 0.f, 0.f, 0.f); else zero = 0.f;

    auto tile_layer = [&](long long tile) {
        int l = 0;
        while (l + 1 < g.L && tile >= p.tile_first[l + 1]) ++l;
        return l;
    };
    // owner code of row `lane` of a tile (-1 beyond the tile / the layer's last row)
    auto load_codes = [&](long long tile) {
        if (tile >= total || (p.debug & 1)) return -1;
        const int l = tile_layer(tile);
        const long long row = (tile - p.tile_first[l]) * p.R[l] + lane;
        if (lane >= p.R[l] || row >= p.rows[l]) return -1;
        return __ldg(a.table + (size_t)a.B * g.cell_off[l] + row);
    };

    int next_codes = load_codes(warp);
    for (long long tile = warp; tile < total; tile += n_warps) {
        const int codes = next_codes;
        next_codes = load_codes(tile + n_warps);          // prefetch: hides the only load
        const int l = tile_layer(tile);
        const int dv = p.dv[l];
        const long long row0 = (tile - p.tile_first[l]) * p.R[l];
        const long long left = p.rows[l] - row0;
        const int n_rows = (int)(left < p.R[l] ? left : p.R[l]);
        const int n_units = n_rows * dv;

        // ---- pass 1: the whole tile as zeros, 32 consecutive units per warp store --------
        V* dst = reinterpret_cast<V*>(a.y[l]) + row0 * dv;
        #pragma unroll 4
        for (int f = lane; f < n_units; f += 32) dst[f] = zero;

        // ---- pass 2: patch the rows that have an owner (7 scalars each) -----------------
        // The lines were just written by this warp and are still in L2, so the patch
        // merges there: DRAM sees every byte once.  Four rows per sweep, 8 lanes each.
        unsigned pos = __ballot_sync(0xffffffffu, codes >= 0);
        if (pos) {
            __syncwarp();                                  // order the patches after the zeros
            const int slot = lane >> 3, role = lane & 7;
            float* ybase = a.y[l] + row0 * (long long)g.D[l];
            while (pos) {
                const unsigned r = nth_bit(pos, slot);
                const int code = __shfl_sync(0xffffffffu, codes, r & 31);
                if (r < 32u && role < 7) {
                    const BoxRec* rec = a.recs + (code >> 4);
                    const int nb = code & 15;
                    int ch = role;
                    float v = 1.0f;
                    if (role == 0)      v = (float)__dadd_rn((double)(1 - nb / 3), rec->fx);   // -dx + fx, :3467
                    else if (role == 1) v = (float)__dadd_rn((double)(1 - nb % 3), rec->fy);   // -dy + fy
                    else if (role == 2) v = rec->tw;
                    else if (role == 3) v = rec->th;
                    else if (role == 5) ch = rec->hot_anchor;                                  // :3469
                    else if (role == 6) ch = rec->hot_class;                                   // :3470
                    ybase[(long long)r * g.D[l] + ch] = v;
                }
                #pragma unroll
                for (int q = 0; q < 4; ++q) pos &= pos - 1;
            }
        }
    }
}

}  // namespace

size_t encode_assign_smem_bytes(const HeadGeom& g, int N)
{
    return (size_t)(2 * g.cells + N) * sizeof(int);
}

bool encode_needs_big_tables(const HeadGeom& g, int N)
{
    return encode_assign_smem_bytes(g, N) > 200 * 1024;
}

cudaError_t launch_encode(const EncodeArgs& a, int num_sms, cudaStream_t stream)
{
    cudaError_t err = launch_encode_assign(a, stream);
    if (err != cudaSuccess) return err;
    return launch_encode_fill(a, num_sms, stream);
}

cudaError_t launch_encode_assign(const EncodeArgs& a, cudaStream_t stream, bool overlap_previous)
{
    const HeadGeom& g = a.g;
    const size_t smem = a.big_tables ? (size_t)a.N * sizeof(int) + 16 : encode_assign_smem_bytes(g, a.N);
    cudaError_t err = cudaFuncSetAttribute(encode_assign_kernel,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) return err;
    prof_mark_begin(PROF_ENCODE_ASSIGN, stream);
    if (overlap_previous) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)a.B);
        cfg.blockDim = dim3(kAssignThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr.val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        const cudaError_t e = cudaLaunchKernelEx(&cfg, encode_assign_kernel, a);
        prof_mark_end(PROF_ENCODE_ASSIGN, stream);
        return e != cudaSuccess ? e : cudaGetLastError();
    }
    encode_assign_kernel<<<a.B, kAssignThreads, smem, stream>>>(a);
    prof_mark_end(PROF_ENCODE_ASSIGN, stream);
    return cudaGetLastError();
}

cudaError_t launch_encode_fill(const EncodeArgs& a, int num_sms, cudaStream_t stream, bool overlap_previous)
{
    const HeadGeom& g = a.g;
    bool vec4 = true;
    for (int l = 0; l < g.L; ++l)
        vec4 = vec4 && (g.D[l] % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.y[l]) & 15) == 0);
    const int vec = vec4 ? 4 : 1;
    FillPlan p;
    long long tiles = 0;
    for (int l = 0; l < g.L; ++l) {
        p.dv[l] = g.D[l] / vec;
        int gcd = p.dv[l], t = 32;
        while (t) { const int m = gcd % t; gcd = t; t = m; }
        p.R[l] = 32 / gcd;
        p.rows[l] = (long long)a.B * g.gh[l] * g.gw[l];
        p.tile_first[l] = tiles;
        tiles += (p.rows[l] + p.R[l] - 1) / p.R[l];
    }
    p.tile_first[g.L] = tiles;
    static int env_dbg = -1;
    if (env_dbg < 0) { const char* e = getenv("MGD_FILL_DEBUG"); env_dbg = e ? atoi(e) : 0; }
    p.debug = env_dbg;
    // One tile per warp and no persistent loop: on B200 a write-only stream reaches
    // ~7.5 TB/s with hundreds of thousands of short-lived CTAs but only ~6 TB/s from a
    // single resident wave looping over the same bytes (scripts/probes/fill_probe.cu).
    long long blocks = (tiles * 32 + kFillThreads - 1) / kFillThreads;
    static int env_waves = -1;
    if (env_waves < 0) { const char* e = getenv("MGD_FILL_MAX_WAVES"); env_waves = e ? atoi(e) : 0; }
    if (env_waves > 0) {
        const long long max_blocks = (long long)num_sms * (2048 / kFillThreads) * env_waves;
        if (blocks > max_blocks) blocks = max_blocks;
    }
    if (blocks > 0x7fffffffll) blocks = 0x7fffffffll;
    if (blocks < 1) blocks = 1;
    prof_mark_begin(PROF_ENCODE_FILL, stream);
    if (overlap_previous) {
        // (only where the previous kernel on `stream` is the writer of another chunk whose
        //  tables are still alive: mgd_encode_decode_nms)
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)blocks);
        cfg.blockDim = dim3(kFillThreads);
        cfg.stream = stream;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr.val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        const cudaError_t e = vec4 ? cudaLaunchKernelEx(&cfg, encode_fill_kernel<4>, a, p)
                                   : cudaLaunchKernelEx(&cfg, encode_fill_kernel<1>, a, p);
        prof_mark_end(PROF_ENCODE_FILL, stream);
        return e != cudaSuccess ? e : cudaGetLastError();
    }
    if (vec4) encode_fill_kernel<4><<<(unsigned)blocks, kFillThreads, 0, stream>>>(a, p);
    else      encode_fill_kernel<1><<<(unsigned)blocks, kFillThreads, 0, stream>>>(a, p);
    prof_mark_end(PROF_ENCODE_FILL, stream);
    return cudaGetLastError();
}
