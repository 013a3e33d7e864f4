// HungarianMatcher on the GPU: batched cost matrix + rectangular linear-sum assignment.
//
// Replaces detr/matcher.py:40-99 (per-image Python loop, ~25 ATen kernels, `.cpu()` sync and one SciPy
// call per image per decoder layer) by ONE launch over all (image, layer) problems:
//   phase 1  (all 4 warps)  cost matrix into shared memory: warp-level softmax statistics, boxes staged in
//                           shared memory, C = w_bbox*L1 + w_class*(-p[label]) + w_giou*(-GIoU)
//                           (detr/matcher.py:66-93, detr/utils.py:57-97), fp32 with the reference's operation
//                           order (explicit _rn intrinsics: no FMA contraction).
//   phase 2  (warp 0)       shortest-augmenting-path assignment in float64 that reproduces SciPy's
//                           linear_sum_assignment tie-breaking bit for bit (oracle/lsap.c):
//                           each lane owns columns j = lane + 32k with their dual v, tentative distance and
//                           scan position in registers; the per-iteration arg-min with SciPy's "array order /
//                           unassigned wins" rule is three 32-bit warp reductions over (sortable fp64 hi, lo, key).
//
// Work matrix: W[i][j], nr = min(Q,M) rows, nc = max(Q,M) columns; when M < Q SciPy solves the transpose,
// so rows are GT boxes and columns are queries.
#include <math_constants.h>

#include "common.cuh"

namespace detr {

constexpr int kMatchThreads = 128;

struct MatchParams {
    // predictions
    const float* logits; int64_t lg_sb, lg_sl, lg_sq;
    const float* boxes;  int64_t bx_sb, bx_sl, bx_sq;
    // packed targets
    const int64_t* gt_labels; const float* gt_boxes; const int32_t* gt_off; const int32_t* match_off;
    int B, L, Q, K, max_m;
    float w_class, w_bbox, w_giou;
    float* cost_out;          // optional export (required when the work matrix does not fit shared memory)
    int64_t* idx_q; int64_t* idx_gt;   // null => cost only
    int32_t* status;
    const int32_t* order;     // optional: images sorted by descending GT count (largest assignment problems first)
};

struct LsapParams {
    const void* cost; const int64_t* cost_off; const int32_t* nr; const int32_t* nc;
    int n_problems, max_nr, max_nc;
    const int64_t* out_off; int64_t* rows_out; int64_t* cols_out; int32_t* status;
};

// ---- shared-memory carve-up (host and device agree through these helpers) -------------------------------
struct SmemPlan {
    int64_t w_bytes, u_off, ints_off, p1_off, total;
};
__host__ __device__ inline SmemPlan plan_smem(int rows_max, int cols_max, int64_t w_elems, int elem, int Q, int max_m,
                                              bool w_in_smem) {
    SmemPlan s;
    s.w_bytes = w_in_smem ? ((w_elems * elem + 15) / 16) * 16 : 0;
    s.u_off = s.w_bytes;
    s.ints_off = s.u_off + (int64_t)rows_max * 8;
    int64_t n_ints = 3 * (int64_t)cols_max + rows_max + 4;
    s.p1_off = ((s.ints_off + n_ints * 4 + 15) / 16) * 16;
    // phase-1 scratch: per-query 9 floats + (max,sum) 2 floats ; per-gt 9 floats + label int
    s.total = s.p1_off + ((int64_t)Q * 11 + (int64_t)max_m * 10) * 4;
    return s;
}

constexpr int64_t kSmemBudget = 200 * 1024;

// ---- float64 <-> order-preserving uint64 ----------------------------------------------------------------
__device__ __forceinline__ unsigned long long to_sortable(double d) {
    long long b = __double_as_longlong(d + 0.0);  // +0.0 folds -0.0 into +0.0 so that == matches bit equality
    return b < 0 ? ~(unsigned long long)b : ((unsigned long long)b | 0x8000000000000000ull);
}
__device__ __forceinline__ double from_sortable(unsigned long long s) {
    unsigned long long b = (s & 0x8000000000000000ull) ? (s & 0x7fffffffffffffffull) : ~s;
    return __longlong_as_double((long long)b);
}

struct LsapScratch {
    double* u;       // [nr]
    int* pred;       // [nc] predecessor row on the shortest path
    int* row_of_col; // [nc]
    int* col_of_row; // [nr]
    int* todo;       // [nc] unvisited columns, SciPy's `remaining`
    int spare;       // index (relative to pred) of a spare word: target of predecessor stores that must not land
};

// One warp solves one problem.  W(i,j) = W[i*si + j*sj]; UNIT: sj == 1 (the shared-memory work matrix).  Returns 0 or a
// DETR_ST_* bit.
//
// The scan of one row is a chain of ~140 dependent-ish instructions issued by a single warp (~3 cycles each), so the
// instruction count IS the run time (ncu: 203 instructions and ~600 cycles per scan before this layout).  Hence:
//  * the lane-local arg-min over the SLOTS columns runs on doubles (DSETP), only the winner is mapped to its sortable bits;
//  * the tie-break key is one IMAD per column: key = kb0 + ksg * pos, with (kb0, ksg) fixed while a row is being inserted;
//  * loads use a clamped column index fixed per problem instead of a per-scan select;
//  * the set of removed columns is read off `pos` after the scan loop instead of being tracked inside it.
template <int SLOTS, typename T, bool UNIT>
__device__ int lsap_solve_warp(const T* __restrict__ W, int64_t si64, int64_t sj64, int nr, int nc, LsapScratch s, int lane) {
    const int si = (int)si64, sj = UNIT ? 1 : (int)sj64;   // nr, nc <= 1024: every element offset fits 32 bits
    double dist[SLOTS], v[SLOTS];
    int rowof[SLOTS], pos[SLOTS], ksg[SLOTS];
    unsigned kb0[SLOTS];
#pragma unroll
    for (int k = 0; k < SLOTS; ++k) { v[k] = 0.0; rowof[k] = -1; dist[k] = 0.0; pos[k] = -1; }
    for (int i = lane; i < nr; i += 32) { s.u[i] = 0.0; s.col_of_row[i] = -1; }
    for (int j = lane; j < nc; j += 32) { s.row_of_col[j] = -1; s.pred[j] = -1; }
    __syncwarp();

    const unsigned long long INF_S = to_sortable(CUDART_INF);

    for (int cur = 0; cur < nr; ++cur) {
#pragma unroll
        for (int k = 0; k < SLOTS; ++k) {
            const int j = lane + 32 * k;
            dist[k] = CUDART_INF;
            pos[k] = (j < nc) ? (nc - 1 - j) : -1;  // todo[t] = nc-1-t  (reverse fill)
            // smaller key wins among equal distances: unassigned columns first, LAST in array order;
            // otherwise assigned columns, FIRST in array order.  key = flag | order | column | row_of_col
            //   unassigned: ((1023 - pos) << 20) | (j << 10)              = kb0 - (pos << 20)
            //   assigned:   (1 << 30) | (pos << 20) | (j << 10) | rowof   = kb0 + (pos << 20)
            kb0[k] = rowof[k] < 0 ? ((1023u << 20) | ((unsigned)j << 10)) : ((1u << 30) | ((unsigned)j << 10) | (unsigned)rowof[k]);
            ksg[k] = rowof[k] < 0 ? -(1 << 20) : (1 << 20);
        }
        for (int t = lane; t < nc; t += 32) s.todo[t] = nc - 1 - t;
        __syncwarp();

        int n_todo = nc;
        double reach = 0.0;
        int i = cur;
        int sink = -1;

        while (true) {
            const double ui = s.u[i];
            const T* Wi = W + i * si;
            // all loads first, then the float64 chains of the SLOTS columns interleaved (branch-free)
            T wv[SLOTS];
#pragma unroll
            for (int k = 0; k < SLOTS; ++k) wv[k] = Wi[min(lane + 32 * k, nc - 1) * sj];
            // Every chain is evaluated unconditionally and kept out of any branch (the empty asm pins r[k] here): when the
            // compiler guards a chain by "column still active" the SLOTS chains end up in SLOTS divergent regions, one after
            // the other (measured: 1.06 -> 1.27 ms at config 3).  Improvements are applied by selects; the predecessor store
            // of a column that does not improve goes to a spare word.
            double r[SLOTS];
#pragma unroll
            for (int k = 0; k < SLOTS; ++k) {
                // ((reach + c) - u_i) - v_j : SciPy's evaluation order, float64
                r[k] = __dsub_rn(__dsub_rn(__dadd_rn(reach, (double)wv[k]), ui), v[k]);
                asm volatile("" : "+d"(r[k]));
            }
            double cd[SLOTS];
            unsigned ck[SLOTS];
#pragma unroll
            for (int k = 0; k < SLOTS; ++k) {
                const bool act = pos[k] >= 0;
                const bool imp = act & (r[k] < dist[k]);
                dist[k] = imp ? r[k] : dist[k];
                s.pred[imp ? lane + 32 * k : s.spare] = i;
                cd[k] = act ? dist[k] : CUDART_INF;
                ck[k] = act ? kb0[k] + (unsigned)(ksg[k] * pos[k]) : ~0u;
            }
            // lane-local arg-min as a balanced tree ((distance, key) is a total order, so the grouping is free): two levels
            // of dependent compares for 4 columns instead of three.  -0.0 == +0.0 here, as in SciPy's comparisons.
#pragma unroll
            for (int w = 1; w < SLOTS; w *= 2) {
#pragma unroll
                for (int k = 0; k + w < SLOTS; k += 2 * w) {
                    const bool better = cd[k + w] < cd[k] || (cd[k + w] == cd[k] && ck[k + w] < ck[k]);
                    cd[k] = better ? cd[k + w] : cd[k];
                    ck[k] = better ? ck[k + w] : ck[k];
                }
            }
            const unsigned long long best = to_sortable(cd[0]);
            const unsigned hi = (unsigned)(best >> 32), lo = (unsigned)best;
            const unsigned mhi = __reduce_min_sync(FULL_MASK, hi);
#ifdef DETR_LSAP_BALLOT
            // single lane with the minimal high word: its (low word, key) by two independent shuffles (measured slower than
            // the two dependent reductions: 0.84 -> 0.91 ms at config 3)
            const unsigned cand = __ballot_sync(FULL_MASK, hi == mhi);
            unsigned mlo, mkey;
            if ((cand & (cand - 1u)) == 0u) {   // warp-uniform
                const int src = __ffs((int)cand) - 1;
                mlo = __shfl_sync(FULL_MASK, lo, src);
                mkey = __shfl_sync(FULL_MASK, ck[0], src);
            } else {
                mlo = __reduce_min_sync(FULL_MASK, hi == mhi ? lo : 0xffffffffu);
                mkey = __reduce_min_sync(FULL_MASK, (hi == mhi && lo == mlo) ? ck[0] : 0xffffffffu);
            }
#else
            const unsigned mlo = __reduce_min_sync(FULL_MASK, hi == mhi ? lo : 0xffffffffu);
            const unsigned mkey = __reduce_min_sync(FULL_MASK, (hi == mhi && lo == mlo) ? ck[0] : 0xffffffffu);
#endif
            const unsigned long long mbest = ((unsigned long long)mhi << 32) | mlo;
            if (mbest >= INF_S) return DETR_ST_INFEASIBLE;
            reach = from_sortable(mbest);

            const bool assigned = (mkey >> 30) & 1u;
            const int jstar = (int)((mkey >> 10) & 1023u);
            const int ord = (int)((mkey >> 20) & 1023u);
            const int tstar = assigned ? ord : 1023 - ord;

            // swap-remove.  Every lane performs the same store, so each lane later reads its own write:
            // no warp barrier is needed on this serial critical path.
            const int jlast = s.todo[n_todo - 1];
            s.todo[tstar] = jlast;
            --n_todo;
#pragma unroll
            for (int k = 0; k < SLOTS; ++k) {
                const int j = lane + 32 * k;
                pos[k] = j == jlast ? tstar : pos[k];
                pos[k] = j == jstar ? -1 : pos[k];   // (jstar == jlast when the last entry is the one removed)
            }
            if (!assigned) { sink = jstar; break; }
            i = (int)(mkey & 1023u);
        }

        // dual update (rows reached are exactly row_of_col of the removed, non-sink columns, plus `cur`); a column j < nc
        // without a position has been removed during this scan
#pragma unroll
        for (int k = 0; k < SLOTS; ++k) {
            const int j = lane + 32 * k;
            if (j < nc && pos[k] < 0) {
                const double delta = __dsub_rn(reach, dist[k]);
                if (j != sink) s.u[rowof[k]] = __dadd_rn(s.u[rowof[k]], delta);
                v[k] = __dsub_rn(v[k], delta);
            }
        }
        if (lane == 0) s.u[cur] = __dadd_rn(s.u[cur], reach);
        __syncwarp();
        // augment along the predecessor chain
        if (lane == 0) {
            int j = sink;
            while (true) {
                const int a = s.pred[j];
                s.row_of_col[j] = a;
                const int prev = s.col_of_row[a];
                s.col_of_row[a] = j;
                j = prev;
                if (a == cur) break;
            }
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < SLOTS; ++k) {
            const int j = lane + 32 * k;
            if (j < nc) rowof[k] = s.row_of_col[j];
        }
    }
    return 0;
}

// Emit pairs sorted by ORIGINAL row (SciPy's output contract).  orig rows = queries for the matcher.
__device__ void emit_pairs(bool flip, int nr, int nc, const LsapScratch& s, int64_t* rows_out, int64_t* cols_out, int lane) {
    if (!flip) {
        for (int i = lane; i < nr; i += 32) { rows_out[i] = i; cols_out[i] = s.col_of_row[i]; }
    } else {
        // working columns are original rows; compact those that are matched, ascending
        int base = 0;
        for (int j0 = 0; j0 < nc; j0 += 32) {
            const int j = j0 + lane;
            const int r = (j < nc) ? s.row_of_col[j] : -1;
            const unsigned m = __ballot_sync(FULL_MASK, r >= 0);
            if (r >= 0) {
                const int k = base + __popc(m & ((1u << lane) - 1u));
                rows_out[k] = j;
                cols_out[k] = r;
            }
            base += __popc(m);
        }
    }
}

__device__ __forceinline__ LsapScratch carve(char* smem, const SmemPlan& pl, int rows_max, int cols_max) {
    LsapScratch s;
    s.u = reinterpret_cast<double*>(smem + pl.u_off);
    int* ip = reinterpret_cast<int*>(smem + pl.ints_off);
    s.pred = ip;
    s.row_of_col = ip + cols_max;
    s.todo = ip + 2 * cols_max;
    s.col_of_row = ip + 3 * cols_max;
    s.spare = 3 * cols_max + rows_max;   // first of the 4 spare ints of plan_smem's n_ints
    return s;
}

// =========================================================================================================
// Fused matcher kernel: one CTA per (image, layer).
// =========================================================================================================
template <int SLOTS, bool W_SMEM>
__global__ void __launch_bounds__(kMatchThreads) hungarian_match_kernel(const MatchParams p) {
    extern __shared__ __align__(16) char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int bi = blockIdx.x / p.L, l = blockIdx.x % p.L;
    const int b = p.order ? p.order[bi] : bi;
    const int g0 = p.gt_off[b];
    const int M = p.gt_off[b + 1] - g0;
    if (M == 0) return;
    const int Q = p.Q, K = p.K;
    const bool flip = M < Q;
    const int nr = flip ? M : Q, nc = flip ? Q : M;
    const int rows_max = min(Q, p.max_m), cols_max = max(Q, p.max_m);
    const SmemPlan pl = plan_smem(rows_max, cols_max, (int64_t)Q * p.max_m, 4, Q, p.max_m, W_SMEM);

    float* pq = reinterpret_cast<float*>(smem + pl.p1_off);  // [Q][11]: cx cy w h x1 y1 x2 y2 area rowmax rowsum
    float* gq = pq + (int64_t)Q * 11;                        // [M][10]: gx1 gy1 gx2 gy2 gcx gcy gw gh area label
    __shared__ int s_flags;
    if (tid == 0) s_flags = 0;

    float* Cg = p.cost_out ? p.cost_out + (int64_t)Q * ((int64_t)p.L * g0 + (int64_t)l * M) : nullptr;
    float* Wsm = reinterpret_cast<float*>(smem);
    // work-matrix addressing
    float* W = W_SMEM ? Wsm : Cg;
    const int64_t si = W_SMEM ? nc : (flip ? 1 : M);
    const int64_t sj = W_SMEM ? 1 : (flip ? M : 1);

    const float* lg = p.logits + b * p.lg_sb + l * p.lg_sl;
    const float* bx = p.boxes + b * p.bx_sb + l * p.bx_sl;
    int local_flags = 0;

    // ---- per-query and per-gt derived quantities (boxes staged in shared memory) ----
    for (int q = tid; q < Q; q += kMatchThreads) {
        const float4 c = *reinterpret_cast<const float4*>(bx + (int64_t)q * p.bx_sq);
        // CXCYWH -> XYXY: x1 = cx - w/2 ; x2 = w + x1   (torchvision _meta.py:158-182)
        const float x1 = __fsub_rn(c.x, __fdiv_rn(c.z, 2.f)), y1 = __fsub_rn(c.y, __fdiv_rn(c.w, 2.f));
        const float x2 = __fadd_rn(c.z, x1), y2 = __fadd_rn(c.w, y1);
        float* o = pq + q * 11;
        o[0] = c.x; o[1] = c.y; o[2] = c.z; o[3] = c.w; o[4] = x1; o[5] = y1; o[6] = x2; o[7] = y2;
        o[8] = __fmul_rn(__fsub_rn(x2, x1), __fsub_rn(y2, y1));
        if (!(x2 >= x1) || !(y2 >= y1)) local_flags |= DETR_ST_DEGENERATE_BOX;
    }
    for (int m = tid; m < M; m += kMatchThreads) {
        const float4 g = *reinterpret_cast<const float4*>(p.gt_boxes + (int64_t)(g0 + m) * 4);
        // XYXY -> CXCYWH: w = x2-x1 ; cx = (2*x1 + w)/2   (torchvision _meta.py:185-194)
        const float gw = __fsub_rn(g.z, g.x), gh = __fsub_rn(g.w, g.y);
        float* o = gq + m * 10;
        o[0] = g.x; o[1] = g.y; o[2] = g.z; o[3] = g.w;
        o[4] = __fdiv_rn(__fadd_rn(__fmul_rn(g.x, 2.f), gw), 2.f);
        o[5] = __fdiv_rn(__fadd_rn(__fmul_rn(g.y, 2.f), gh), 2.f);
        o[6] = gw; o[7] = gh;
        o[8] = __fmul_rn(gw, gh);
        const int64_t lab = p.gt_labels[g0 + m];
        if (lab < 0 || lab >= K) local_flags |= DETR_ST_BAD_LABEL;
        o[9] = __int_as_float((int)min(max(lab, (int64_t)0), (int64_t)K - 1));
        if (!(g.z >= g.x) || !(g.w >= g.y)) local_flags |= DETR_ST_DEGENERATE_BOX;
    }
    // ---- warp-level softmax statistics, one query row per warp (coalesced); the loads of 4 rows are in flight together
    //      (one row at a time was a chain of L2 round trips: ~25 per warp and problem) ----
    constexpr int kStatRows = 4, kStatChunks = 4;   // register path: K <= 128
    if (K <= 32 * kStatChunks) {
        for (int q0 = warp; q0 < Q; q0 += (kMatchThreads / 32) * kStatRows) {
            float x[kStatRows][kStatChunks];
#pragma unroll
            for (int r = 0; r < kStatRows; ++r) {
                const int q = min(q0 + r * (kMatchThreads / 32), Q - 1);
                const float* row = lg + (int64_t)q * p.lg_sq;
#pragma unroll
                for (int c = 0; c < kStatChunks; ++c) {
                    const int k = lane + 32 * c;
                    x[r][c] = k < K ? row[k] : -CUDART_INF_F;
                }
            }
#pragma unroll
            for (int r = 0; r < kStatRows; ++r) {
                const int q = q0 + r * (kMatchThreads / 32);
                float mx = fmaxf(fmaxf(x[r][0], x[r][1]), fmaxf(x[r][2], x[r][3]));
                mx = warp_max(mx);
                float sum = 0.f;
#pragma unroll
                for (int c = 0; c < kStatChunks; ++c)
                    if (lane + 32 * c < K) sum += expf(x[r][c] - mx);   // same terms, same per-lane order as the generic loop
                sum = warp_sum(sum);
                if (lane == 0 && q < Q) { pq[q * 11 + 9] = mx; pq[q * 11 + 10] = sum; }
            }
        }
    } else {
        for (int q = warp; q < Q; q += kMatchThreads / 32) {
            const float* row = lg + (int64_t)q * p.lg_sq;
            float mx = -CUDART_INF_F;
            for (int k = lane; k < K; k += 32) mx = fmaxf(mx, row[k]);
            mx = warp_max(mx);
            float sum = 0.f;
            for (int k = lane; k < K; k += 32) sum += expf(row[k] - mx);
            sum = warp_sum(sum);
            if (lane == 0) { pq[q * 11 + 9] = mx; pq[q * 11 + 10] = sum; }
        }
    }
    __syncthreads();

    // ---- pairwise cost, thread per (query, gt) pair; fastest index follows W's contiguous dimension ----
    const int total = Q * M;
    for (int e = tid; e < total; e += kMatchThreads) {
        int q, m;
        if (flip) { m = e / Q; q = e - m * Q; } else { q = e / M; m = e - q * M; }
        const float* a = pq + q * 11;
        const float* g = gq + m * 10;
        const int lab = __float_as_int(g[9]);
        const float prob = __fdiv_rn(expf(__ldg(lg + (int64_t)q * p.lg_sq + lab) - a[9]), a[10]);
        // L1 over (cx,cy,w,h)
        float l1 = fabsf(__fsub_rn(a[0], g[4]));
        l1 = __fadd_rn(l1, fabsf(__fsub_rn(a[1], g[5])));
        l1 = __fadd_rn(l1, fabsf(__fsub_rn(a[2], g[6])));
        l1 = __fadd_rn(l1, fabsf(__fsub_rn(a[3], g[7])));
        // GIoU (detr/utils.py:57-97): no eps, extents clamped at 0
        const float iw = fmaxf(__fsub_rn(fminf(a[6], g[2]), fmaxf(a[4], g[0])), 0.f);
        const float ih = fmaxf(__fsub_rn(fminf(a[7], g[3]), fmaxf(a[5], g[1])), 0.f);
        const float inter = __fmul_rn(iw, ih);
        const float uni = __fsub_rn(__fadd_rn(a[8], g[8]), inter);
        const float iou = __fdiv_rn(inter, uni);
        const float hw = fmaxf(__fsub_rn(fmaxf(a[6], g[2]), fminf(a[4], g[0])), 0.f);
        const float hh = fmaxf(__fsub_rn(fmaxf(a[7], g[3]), fminf(a[5], g[1])), 0.f);
        const float hull = __fmul_rn(hw, hh);
        const float giou = __fsub_rn(iou, __fdiv_rn(__fsub_rn(hull, uni), hull));
        // C = w_bbox*L1 + w_class*(-p) + w_giou*(-giou), left to right (detr/matcher.py:93)
        const float c = __fadd_rn(__fadd_rn(__fmul_rn(p.w_bbox, l1), __fmul_rn(p.w_class, -prob)), __fmul_rn(p.w_giou, -giou));
        if (!(c == c) || c == -CUDART_INF_F) local_flags |= DETR_ST_INVALID_COST;
        if (W_SMEM) Wsm[e] = c;                       // e == i*nc + j by construction
        else Cg[(int64_t)q * M + m] = c;
    }
    if (local_flags) atomicOr(&s_flags, local_flags);
    __syncthreads();
    const int flags = s_flags;
    if (flags && tid == 0) atomicOr(p.status, flags);

    if (W_SMEM && Cg) {  // optional export, row-major (query, gt)
        for (int e = tid; e < total; e += kMatchThreads) {
            const int q = e / M, m = e - q * M;
            Cg[e] = flip ? Wsm[m * Q + q] : Wsm[e];
        }
    }
    if (!p.idx_q) return;

    const int n = nr;
    int64_t* oq = p.idx_q + (int64_t)p.L * p.match_off[b] + (int64_t)l * n;
    int64_t* og = p.idx_gt + (int64_t)p.L * p.match_off[b] + (int64_t)l * n;
    if (flags & (DETR_ST_INVALID_COST | DETR_ST_BAD_LABEL)) {  // the reference would have raised: poison the output
        for (int k = tid; k < n; k += kMatchThreads) { oq[k] = -1; og[k] = -1; }
        return;
    }
    if (warp != 0) return;
    if (!W_SMEM) __threadfence_block();
    LsapScratch s = carve(smem, pl, rows_max, cols_max);
    const int rc = lsap_solve_warp<SLOTS, float, W_SMEM>(W, si, sj, nr, nc, s, lane);
    if (rc) {
        if (lane == 0) atomicOr(p.status, rc);
        for (int k = lane; k < n; k += 32) { oq[k] = -1; og[k] = -1; }
        return;
    }
    // original rows = queries.  flip: working columns are queries -> (j, row_of_col[j]); else (i, col_of_row[i])
    emit_pairs(flip, nr, nc, s, oq, og, lane);
}

// =========================================================================================================
// Stand-alone batched LSAP on caller-provided cost matrices (scipy.optimize.linear_sum_assignment drop-in)
// =========================================================================================================
template <int SLOTS, typename T, bool W_SMEM>
__global__ void __launch_bounds__(kMatchThreads) lsap_kernel(const LsapParams p) {
    extern __shared__ __align__(16) char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pid = blockIdx.x;
    const int R0 = p.nr[pid], C0 = p.nc[pid];
    const int n = min(R0, C0);
    if (n == 0) return;
    const bool flip = C0 < R0;  // tall matrix: solve the transpose
    const int nr = flip ? C0 : R0, nc = flip ? R0 : C0;
    const int rows_max = min(p.max_nr, p.max_nc), cols_max = max(p.max_nr, p.max_nc);
    const SmemPlan pl = plan_smem(rows_max, cols_max, (int64_t)p.max_nr * p.max_nc, sizeof(T), 0, 0, W_SMEM);
    const T* C = reinterpret_cast<const T*>(p.cost) + p.cost_off[pid];
    T* Wsm = reinterpret_cast<T*>(smem);
    __shared__ int s_flags;
    if (tid == 0) s_flags = 0;
    __syncthreads();
    int bad = 0;
    const int total = R0 * C0;
    for (int e = tid; e < total; e += kMatchThreads) {  // coalesced read of the row-major input
        const T c = C[e];
        if (!(c == c) || c == (T)(-CUDART_INF)) bad = DETR_ST_INVALID_COST;
        if (W_SMEM) {
            const int r = e / C0, k = e - r * C0;
            Wsm[flip ? (k * nc + r) : e] = c;
        }
    }
    if (bad) atomicOr(&s_flags, bad);
    __syncthreads();
    int64_t* ro = p.rows_out + p.out_off[pid];
    int64_t* co = p.cols_out + p.out_off[pid];
    if (s_flags) {
        if (tid == 0) atomicOr(p.status, s_flags);
        for (int k = tid; k < n; k += kMatchThreads) { ro[k] = -1; co[k] = -1; }
        return;
    }
    if (warp != 0) return;
    LsapScratch s = carve(smem, pl, rows_max, cols_max);
    const T* W = W_SMEM ? Wsm : C;
    const int64_t si = W_SMEM ? nc : (flip ? 1 : C0);
    const int64_t sj = W_SMEM ? 1 : (flip ? C0 : 1);
    const int rc = lsap_solve_warp<SLOTS, T, W_SMEM>(W, si, sj, nr, nc, s, lane);
    if (rc) {
        if (lane == 0) atomicOr(p.status, rc);
        for (int k = lane; k < n; k += 32) { ro[k] = -1; co[k] = -1; }
        return;
    }
    emit_pairs(flip, nr, nc, s, ro, co, lane);
}

// ---- host-side dispatch ----------------------------------------------------------------------------------
template <typename KernelT>
static int launch_with_smem(KernelT kern, int grid, int64_t smem, cudaStream_t st, const char* name) {
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("%s: cannot reserve %lld B of shared memory: %s", name, (long long)smem, cudaGetErrorString(e)); return 2; }
    }
    (void)grid; (void)st;
    return 0;
}

#define DISPATCH_SLOTS(cols, MACRO)          \
    if ((cols) <= 128) { MACRO(4) }          \
    else if ((cols) <= 320) { MACRO(10) }    \
    else if ((cols) <= 512) { MACRO(16) }    \
    else { MACRO(32) }

static int run_match(const MatchParams& p, cudaStream_t st) {
    DETR_CHECK_ARG(p.B >= 0 && p.L >= 1 && p.Q >= 1 && p.K >= 1 && p.max_m >= 0, "matcher: bad sizes B=%d L=%d Q=%d K=%d max_m=%d", p.B, p.L, p.Q, p.K, p.max_m);
    DETR_CHECK_ARG(p.Q <= 1024 && p.max_m <= 1024, "matcher: Q=%d / max_m=%d exceed the 1024 limit of the assignment kernel", p.Q, p.max_m);
    DETR_CHECK_ARG(((uintptr_t)p.boxes % 16) == 0 && (p.bx_sb % 4) == 0 && (p.bx_sl % 4) == 0 && (p.bx_sq % 4) == 0, "matcher: pred boxes must be 16-byte aligned rows");
    if (p.B == 0 || p.max_m == 0) return 0;
    const int rows_max = p.Q < p.max_m ? p.Q : p.max_m, cols_max = p.Q < p.max_m ? p.max_m : p.Q;
    SmemPlan pl = plan_smem(rows_max, cols_max, (int64_t)p.Q * p.max_m, 4, p.Q, p.max_m, true);
    const bool in_smem = pl.total <= kSmemBudget;
    if (!in_smem) {
        DETR_CHECK_ARG(p.cost_out != nullptr, "matcher: Q*max_m=%d*%d does not fit shared memory; cost_out workspace required", p.Q, p.max_m);
        pl = plan_smem(rows_max, cols_max, (int64_t)p.Q * p.max_m, 4, p.Q, p.max_m, false);
    }
    const int grid = p.B * p.L;
#define LAUNCH_MATCH(S)                                                                                      \
    if (in_smem) {                                                                                           \
        if (launch_with_smem(hungarian_match_kernel<S, true>, grid, pl.total, st, "hungarian_match")) return 2; \
        hungarian_match_kernel<S, true><<<grid, kMatchThreads, pl.total, st>>>(p);                           \
    } else {                                                                                                 \
        if (launch_with_smem(hungarian_match_kernel<S, false>, grid, pl.total, st, "hungarian_match")) return 2; \
        hungarian_match_kernel<S, false><<<grid, kMatchThreads, pl.total, st>>>(p);                          \
    }
    DISPATCH_SLOTS(cols_max, LAUNCH_MATCH)
#undef LAUNCH_MATCH
    DETR_CHECK_LAUNCH("hungarian_match");
    return 0;
}

template <typename T>
static int run_lsap(const LsapParams& p, cudaStream_t st) {
    DETR_CHECK_ARG(p.n_problems >= 0 && p.max_nr >= 0 && p.max_nc >= 0, "lsap: bad sizes");
    DETR_CHECK_ARG(p.max_nr <= 1024 && p.max_nc <= 1024, "lsap: %d x %d exceeds the 1024 limit of the assignment kernel", p.max_nr, p.max_nc);
    if (p.n_problems == 0 || p.max_nr == 0 || p.max_nc == 0) return 0;
    const int rows_max = p.max_nr < p.max_nc ? p.max_nr : p.max_nc, cols_max = p.max_nr < p.max_nc ? p.max_nc : p.max_nr;
    SmemPlan pl = plan_smem(rows_max, cols_max, (int64_t)p.max_nr * p.max_nc, sizeof(T), 0, 0, true);
    const bool in_smem = pl.total <= kSmemBudget;
    if (!in_smem) pl = plan_smem(rows_max, cols_max, (int64_t)p.max_nr * p.max_nc, sizeof(T), 0, 0, false);
#define LAUNCH_LSAP(S)                                                                                  \
    if (in_smem) {                                                                                      \
        if (launch_with_smem(lsap_kernel<S, T, true>, p.n_problems, pl.total, st, "lsap")) return 2;    \
        lsap_kernel<S, T, true><<<p.n_problems, kMatchThreads, pl.total, st>>>(p);                      \
    } else {                                                                                            \
        if (launch_with_smem(lsap_kernel<S, T, false>, p.n_problems, pl.total, st, "lsap")) return 2;   \
        lsap_kernel<S, T, false><<<p.n_problems, kMatchThreads, pl.total, st>>>(p);                     \
    }
    DISPATCH_SLOTS(cols_max, LAUNCH_LSAP)
#undef LAUNCH_LSAP
    DETR_CHECK_LAUNCH("lsap");
    return 0;
}

}  // namespace detr

using namespace detr;

extern "C" int64_t detr_matcher_smem_bytes(int Q, int max_m, int elem_size) {
    const int rows_max = Q < max_m ? Q : max_m, cols_max = Q < max_m ? max_m : Q;
    SmemPlan pl = plan_smem(rows_max, cols_max, (int64_t)Q * max_m, elem_size, Q, max_m, true);
    return pl.total <= kSmemBudget ? pl.total : -1;
}

extern "C" int detr_cost_matrix_f32(const float* logits, int64_t lg_sb, int64_t lg_sl, int64_t lg_sq,
                                    const float* boxes, int64_t bx_sb, int64_t bx_sl, int64_t bx_sq,
                                    const int64_t* gt_labels, const float* gt_boxes, const int32_t* gt_off,
                                    int B, int L, int Q, int K, int max_m, float w_class, float w_bbox, float w_giou,
                                    float* cost_out, int32_t* status, void* stream) {
    DETR_CHECK_ARG(cost_out != nullptr && status != nullptr, "cost_matrix: cost_out and status are required");
    MatchParams p{logits, lg_sb, lg_sl, lg_sq, boxes, bx_sb, bx_sl, bx_sq, gt_labels, gt_boxes, gt_off, nullptr,
                  B, L, Q, K, max_m, w_class, w_bbox, w_giou, cost_out, nullptr, nullptr, status, nullptr};
    return run_match(p, (cudaStream_t)stream);
}

// order[rank] = image index, rank by descending GT count (ties by index): the assignment phase of a problem costs ~M^2
// serial steps, so a launch with more problems than resident CTAs finishes sooner when the big ones start first.
__global__ void match_order_kernel(const int32_t* __restrict__ gt_off, int B, int32_t* __restrict__ order) {
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        const int m = gt_off[b + 1] - gt_off[b];
        int rank = 0;
        for (int c = 0; c < B; ++c) {
            const int mc = gt_off[c + 1] - gt_off[c];
            rank += (mc > m) || (mc == m && c < b);
        }
        order[rank] = b;
    }
}

extern "C" int detr_hungarian_match_f32(const float* logits, int64_t lg_sb, int64_t lg_sl, int64_t lg_sq,
                                        const float* boxes, int64_t bx_sb, int64_t bx_sl, int64_t bx_sq,
                                        const int64_t* gt_labels, const float* gt_boxes, const int32_t* gt_off,
                                        const int32_t* match_off, int B, int L, int Q, int K, int max_m,
                                        float w_class, float w_bbox, float w_giou, float* cost_out,
                                        int64_t* idx_q, int64_t* idx_gt, int32_t* status, int32_t* order_ws, void* stream) {
    DETR_CHECK_ARG(idx_q != nullptr && idx_gt != nullptr && status != nullptr && match_off != nullptr,
                   "hungarian_match: idx_q, idx_gt, match_off and status are required");
    DETR_CHECK_ARG(!(w_class == 0.f && w_bbox == 0.f && w_giou == 0.f), "all costs can't be 0");  // detr/matcher.py:38
    MatchParams p{logits, lg_sb, lg_sl, lg_sq, boxes, bx_sb, bx_sl, bx_sq, gt_labels, gt_boxes, gt_off, match_off,
                  B, L, Q, K, max_m, w_class, w_bbox, w_giou, cost_out, idx_q, idx_gt, status, nullptr};
    // more problems than CTA slots of one wave (~4 per SM): run the largest first
    if (order_ws != nullptr && (int64_t)B * L > 4 * 148 && B <= 4096) {
        match_order_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(gt_off, B, order_ws);
        DETR_CHECK_LAUNCH("match_order");
        p.order = order_ws;
    }
    return run_match(p, (cudaStream_t)stream);
}

extern "C" int detr_lsap_f32(const float* cost, const int64_t* cost_off, const int32_t* nr, const int32_t* nc,
                             int n_problems, int max_nr, int max_nc, const int64_t* out_off, int64_t* rows_out,
                             int64_t* cols_out, int32_t* status, void* stream) {
    LsapParams p{cost, cost_off, nr, nc, n_problems, max_nr, max_nc, out_off, rows_out, cols_out, status};
    return run_lsap<float>(p, (cudaStream_t)stream);
}

extern "C" int detr_lsap_f64(const double* cost, const int64_t* cost_off, const int32_t* nr, const int32_t* nc,
                             int n_problems, int max_nr, int max_nc, const int64_t* out_off, int64_t* rows_out,
                             int64_t* cols_out, int32_t* status, void* stream) {
    LsapParams p{cost, cost_off, nr, nc, n_problems, max_nr, max_nc, out_off, rows_out, cols_out, status};
    return run_lsap<double>(p, (cudaStream_t)stream);
}
