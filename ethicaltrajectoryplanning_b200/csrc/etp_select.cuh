// Selection of the best candidate, device side shared by etp_select.cu (its own launches) and etp_score.cu (a planning cycle
// runs the selection in the fused kernel's last CTA).  See etp_select.cu for the scheme.
#pragma once

#include "etp_harm.cuh"
#include "etp_launch.h"

namespace etp {

constexpr int kSelThreads = 256;
constexpr int kSelPer = 16;                       // workspace entries per thread
constexpr int kSelChunk = kSelThreads * kSelPer;  // entries per block
constexpr int kResThreads = 512;
constexpr int kQChunk = 128;                      // un-gated pairs whose 36 BVU values are buffered at a time (1.2 KB each)
constexpr int kQChunkBatch = 32;                  // the same in the sweep's kernel (one CTA per scenario: several resident per SM)
constexpr int kEntC = 6;                          // per-pair constants kept in shared memory

struct SelBlock {
    int rank_max;         // highest level rank among the block's unambiguous candidates (-1: none)
    int n_scored, n_amb;  // non-skipped / ambiguous candidates
    int arg;              // column of the smallest (upper, column) at rank_max
    double upper, lower;  // min of cost + bound / of cost - bound at rank_max
};

struct SelState {  // what select_kernel leaves for select_collect_kernel (steps with more than one block)
    double thr;    // smallest upper bound cost + bound of the top level
    int hi;        // top level rank among the unambiguous candidates
    int n_scored, n_amb;
    int pending;   // 1: the survivors still have to be collected
};

struct RescoreRec {
    double cost;
    int rank, level, cl, pad;
};

__device__ __forceinline__ double order_cost(double c) { return c != c ? INFINITY : c; }  // NaN costs never win

__device__ __forceinline__ int dense_of(const DevStep& st, bool given, int cl) {
    if (given) return cl;
    const int kl = cl / st.nd, id = cl - kl * st.nd;
    return (st.p_begin + st.shard_index + kl * st.shard_count) * st.nd + id;
}

// deterministic block sum (fixed tree): every thread gets the total
__device__ __forceinline__ double block_sum(double v, double* scratch) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += scratch[w];
    return s;
}

// the same for several values at once (one pair of barriers for all of them); scratch holds N x (threads / 32) doubles
template <int N> __device__ __forceinline__ void block_sum_n(double (&v)[N], double* scratch) {
    const int nw = (int)(blockDim.x >> 5);
#pragma unroll
    for (int k = 0; k < N; ++k)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(kFull, v[k], o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) scratch[k * nw + (threadIdx.x >> 5)] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < N; ++k) {
        double s = 0.0;
        for (int w = 0; w < nw; ++w) s += scratch[k * nw + w];
        v[k] = s;
    }
}

// rank of dense candidate `index` among the non-skipped candidates of [c_begin, c_end) = position in the reference's
// fp_list; all threads of the block take part
__device__ inline int list_index_of(const DevStep& st, const int* __restrict__ nskip, bool given, int index, double* scratch) {
    if (given || index < 0) return index;
    const int pb = index / st.nd, idb = index % st.nd;
    double n = 0.0;
    for (int pp = st.p_begin + threadIdx.x; pp < pb; pp += blockDim.x) n += (double)(st.nd - nskip[pp - st.p_begin]);
    const double dsb = st.prof_meta[(size_t)(pb - st.p_begin) * PM_COUNT + PM_DS];
    for (int id = threadIdx.x; id < idb; id += blockDim.x) n += (dsb <= fabs(st.d_list[id])) ? 0.0 : 1.0;
    return (int)block_sum(n, scratch);
}

__device__ __forceinline__ void write_best(BestRec* out, int level, double cost, int index, int n_scored, int list_index) {
    BestRec b;
    b.key_hi = index < 0 ? ~0ull : make_key_hi(level, cost);
    b.key_lo = index < 0 ? ~0ull : make_key_lo(cost, (unsigned long long)index);
    b.cost = cost;
    b.index = index;
    b.level = level;
    b.n_scored = n_scored;
    b.list_index = list_index;
    *out = b;
}

// Survivors that are one and the same computation: the current velocity is appended to the linspace of end velocities
// (frenet_functions.py:760-807) and coincides with a sample whenever the ego drives at a limit of the range, so two profiles
// -- and with them whole rows of candidates -- are bit-identical in every input.  Whatever their exact cost, the first in
// generation order wins (stable sort, frenet_planner.py:457).  Returns that column if every survivor duplicates it, else -1.
// One thread; `count` is small.
__device__ inline int duplicate_winner(const DevStep& st, const volatile int* list, int count) {
    if (st.ext_level || st.bd_harm || st.factor_list) return -1;  // per-candidate inputs of the host
    int m = list[0];
    for (int i = 1; i < count; ++i) m = min(m, list[i]);
    const int klm = m / st.nd, idm = m - klm * st.nd;
    const int pm = st.p_begin + st.shard_index + klm * st.shard_count;
    const double vm = st.v_list[pm / st.nt], tm = st.t_list[pm % st.nt];
    for (int i = 0; i < count; ++i) {
        const int cl = list[i];
        if (cl == m) continue;
        const int kl = cl / st.nd, id = cl - kl * st.nd;
        const int p = st.p_begin + st.shard_index + kl * st.shard_count;
        if (id != idm || st.v_list[p / st.nt] != vm || st.t_list[p % st.nt] != tm) return -1;
    }
    return m;
}

// ---------------------------------------------------------------------------------------------
// select_kernel
// ---------------------------------------------------------------------------------------------
// `blk` of `nblk` blocks work on the step `st` (the whole grid for one planning step; one block per scenario in a sweep)
// With more than one block the survivors are collected by a second launch (select_collect_kernel, every block its own
// chunk) instead of by the last block alone.
__device__ __forceinline__ void select_body(const DevStep& st, const SelectBufs& sb, BestRec* __restrict__ best_out,
                                            const int* __restrict__ nskip, unsigned* __restrict__ tile_counter, int given,
                                            int always_rescore, const int blk, const int nblk) {
    __shared__ int s_i[kSelThreads / 32];
    __shared__ double s_u[kSelThreads / 32], s_l[kSelThreads / 32];
    __shared__ int s_a[kSelThreads / 32];
    __shared__ int s_rank, s_last, s_cnt, s_nsc, s_namb, s_dup;
    __shared__ double s_scratch[kSelThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Cs = st.n_local;
    const bool refine = st.refine != 0;
    SelBlock* blocks = reinterpret_cast<SelBlock*>(sb.blocks);
    const int base = blk * kSelChunk;

    // ---- pass 1: this block's entries ------------------------------------------------------------
    unsigned char meta[kSelPer];
    double lo[kSelPer], up[kSelPer];
    int rmax = -1, nsc = 0, namb = 0;
#pragma unroll
    for (int k = 0; k < kSelPer; ++k) {
        const int cl = base + k * kSelThreads + tid;
        meta[k] = SEL_SKIPPED;
        lo[k] = up[k] = INFINITY;
        if (cl < Cs) {
            meta[k] = st.ws_meta[cl];
            if (meta[k] != SEL_SKIPPED) {
                const double c = order_cost(st.ws_cost[cl]);
                const double b = refine ? (double)st.ws_bound[cl] : 0.0;
                lo[k] = c - b;
                up[k] = c + b;
                ++nsc;
                if (meta[k] & SEL_AMB) {
                    ++namb;
                    sb.list[atomicAdd(&sb.hdr[2], 1)] = cl;  // re-scored whatever the threshold turns out to be
                } else {
                    rmax = max(rmax, level_rank(meta[k]));
                }
            }
        }
    }
    rmax = __reduce_max_sync(kFull, rmax);
    nsc = __reduce_add_sync(kFull, nsc);
    namb = __reduce_add_sync(kFull, namb);
    if (lane == 0) { s_i[warp] = rmax; s_a[warp] = nsc; s_u[warp] = (double)namb; }
    __syncthreads();
    if (tid == 0) {
        int r = -1, a = 0, m = 0;
        for (int w = 0; w < kSelThreads / 32; ++w) { r = max(r, s_i[w]); a += s_a[w]; m += (int)s_u[w]; }
        s_rank = r; s_nsc = a; s_namb = m;
    }
    __syncthreads();
    const int brank = s_rank;
    double bu = INFINITY, bl = INFINITY;
    int ba = 0x7fffffff;
#pragma unroll
    for (int k = 0; k < kSelPer; ++k) {
        if (meta[k] == SEL_SKIPPED || (meta[k] & SEL_AMB) || level_rank(meta[k]) != brank) continue;
        const int cl = base + k * kSelThreads + tid;
        if (up[k] < bu || (up[k] == bu && cl < ba)) { bu = up[k]; ba = cl; }
        bl = fmin(bl, lo[k]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ou = __shfl_xor_sync(kFull, bu, o), ol = __shfl_xor_sync(kFull, bl, o);
        const int oa = __shfl_xor_sync(kFull, ba, o);
        if (ou < bu || (ou == bu && oa < ba)) { bu = ou; ba = oa; }
        bl = fmin(bl, ol);
    }
    __syncthreads();
    if (lane == 0) { s_u[warp] = bu; s_l[warp] = bl; s_a[warp] = ba; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < kSelThreads / 32; ++w) {
            if (s_u[w] < bu || (s_u[w] == bu && s_a[w] < ba)) { bu = s_u[w]; ba = s_a[w]; }
            bl = fmin(bl, s_l[w]);
        }
        SelBlock r;
        r.rank_max = brank; r.n_scored = s_nsc; r.n_amb = s_namb; r.arg = ba; r.upper = bu; r.lower = bl;
        blocks[blk] = r;
        __threadfence();
        s_last = atomicAdd(&sb.counters[0], 1u) == (unsigned)(nblk - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();

    // ---- last block: global top level and threshold ----------------------------------------------
    int hi = -1, tot = 0, amb = 0;
    for (int b = tid; b < nblk; b += kSelThreads) {
        const volatile SelBlock* r = blocks + b;
        hi = max(hi, r->rank_max);
        tot += r->n_scored;
        amb += r->n_amb;
    }
    hi = __reduce_max_sync(kFull, hi);
    tot = __reduce_add_sync(kFull, tot);
    amb = __reduce_add_sync(kFull, amb);
    if (lane == 0) { s_i[warp] = hi; s_a[warp] = tot; s_u[warp] = (double)amb; }
    __syncthreads();
    if (tid == 0) {
        int r = -1, a = 0, m = 0;
        for (int w = 0; w < kSelThreads / 32; ++w) { r = max(r, s_i[w]); a += s_a[w]; m += (int)s_u[w]; }
        s_rank = r; s_nsc = a; s_namb = m; s_cnt = m;  // the list already holds the ambiguous candidates
    }
    __syncthreads();
    hi = s_rank;
    const int n_scored = s_nsc, n_amb = s_namb;
    double thr = INFINITY;
    int arg = 0x7fffffff;
    for (int b = tid; b < nblk; b += kSelThreads) {
        const volatile SelBlock* r = blocks + b;
        if (r->rank_max != hi || hi < 0) continue;
        const double u = r->upper;
        const int a = r->arg;
        if (u < thr || (u == thr && a < arg)) { thr = u; arg = a; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ou = __shfl_xor_sync(kFull, thr, o);
        const int oa = __shfl_xor_sync(kFull, arg, o);
        if (ou < thr || (ou == thr && oa < arg)) { thr = ou; arg = oa; }
    }
    __syncthreads();
    if (lane == 0) { s_u[warp] = thr; s_a[warp] = arg; }
    __syncthreads();
    for (int w = 0; w < kSelThreads / 32; ++w)
        if (s_u[w] < thr || (s_u[w] == thr && s_a[w] < arg)) { thr = s_u[w]; arg = s_a[w]; }
    __syncthreads();

    int done_cl = -2;  // >= -1: the selection is final (column, -1 = no candidate)
    if (n_scored == 0) done_cl = -1;
    else if (!refine) done_cl = arg;  // exact costs: the minimum is the selection (n_amb == 0, so hi >= 0)
    if (nblk > 1) {
        SelState* state = reinterpret_cast<SelState*>(sb.state);
        if (tid == 0) {
            state->thr = thr; state->hi = hi; state->n_scored = n_scored; state->n_amb = n_amb;
            state->pending = done_cl == -2 ? 1 : 0;
            sb.counters[0] = 0;  // re-armed for select_collect_kernel
        }
        if (done_cl == -2) return;  // select_collect_kernel takes it from here
    }
    if (done_cl == -2) {
        // ---- candidates that can still win: ambiguous ones, and the top level's within reach of the threshold --------
        for (int b = 0; b < nblk; ++b) {
            const volatile SelBlock* r = blocks + b;
            const bool look = hi >= 0 && r->rank_max == hi && r->lower <= thr;
            if (!look) continue;
            const int bb = b * kSelChunk;
            for (int k = 0; k < kSelPer; ++k) {
                const int cl = bb + k * kSelThreads + tid;
                if (cl >= Cs) continue;
                const unsigned char m = st.ws_meta[cl];
                if (m == SEL_SKIPPED || (m & SEL_AMB) || level_rank(m) != hi) continue;
                if (order_cost(st.ws_cost[cl]) - (double)st.ws_bound[cl] <= thr) sb.list[atomicAdd(&s_cnt, 1)] = cl;
            }
        }
        __syncthreads();
        if (s_cnt == 1 && n_amb == 0 && !always_rescore) done_cl = sb.list[0];
        else if (s_cnt > 1 && s_cnt <= 64 && n_amb == 0 && !always_rescore && !given) {
            if (tid == 0) s_dup = duplicate_winner(st, sb.list, s_cnt);
            __syncthreads();
            if (s_dup >= 0) done_cl = s_dup;
        }
    }
    if (done_cl != -2) {
        const int index = done_cl < 0 ? -1 : dense_of(st, given != 0, done_cl);
        const int li = list_index_of(st, nskip, given != 0, index, s_scratch);
        if (tid == 0) {
            const int level = done_cl < 0 ? -1 : (int)(st.ws_meta[done_cl] & 63);
            write_best(best_out, level, done_cl < 0 ? 0.0 : st.ws_cost[done_cl], index, n_scored, li);
        }
    }
    if (tid == 0) {
        sb.hdr[0] = done_cl != -2 ? 0 : s_cnt;  // candidates left for rescore_kernel
        sb.hdr[1] = n_scored;
        sb.hdr[2] = 0;
        sb.counters[0] = 0;  // re-arm
        sb.counters[1] = 0;
        if (tile_counter) *tile_counter = 0;  // the fused kernel's tile counter
    }
}

// Second half of the selection of a step with several blocks: every block whose record says it can hold a survivor scans
// its own chunk; the last block to finish closes the list (one survivor, nothing ambiguous, no other shard: the winner).
__device__ __forceinline__ void select_collect_body(const DevStep& st, const SelectBufs& sb, BestRec* __restrict__ best_out,
                                                    const int* __restrict__ nskip, unsigned* __restrict__ tile_counter, int given,
                                                    int always_rescore, const int blk, const int nblk) {
    __shared__ int s_last;
    __shared__ double s_scratch[kSelThreads / 32];
    const SelState* state = reinterpret_cast<const SelState*>(sb.state);
    if (!state->pending) return;  // select_kernel had the answer already (and has re-armed everything)
    const int tid = threadIdx.x, Cs = st.n_local, hi = state->hi;
    const double thr = state->thr;
    const SelBlock* blocks = reinterpret_cast<const SelBlock*>(sb.blocks);
    const SelBlock mine = blocks[blk];
    if (hi >= 0 && mine.rank_max == hi && mine.lower <= thr) {
        const int bb = blk * kSelChunk;
        for (int k = 0; k < kSelPer; ++k) {
            const int cl = bb + k * kSelThreads + tid;
            if (cl >= Cs) continue;
            const unsigned char m = st.ws_meta[cl];
            if (m == SEL_SKIPPED || (m & SEL_AMB) || level_rank(m) != hi) continue;
            if (order_cost(st.ws_cost[cl]) - (double)st.ws_bound[cl] <= thr) sb.list[atomicAdd(&sb.hdr[2], 1)] = cl;
        }
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        s_last = atomicAdd(&sb.counters[0], 1u) == (unsigned)(nblk - 1);
    }
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const int count = *reinterpret_cast<volatile int*>(&sb.hdr[2]);
    bool final_one = count == 1 && state->n_amb == 0 && !always_rescore;
    int cl_one = final_one ? *reinterpret_cast<volatile int*>(&sb.list[0]) : -1;
    if (count > 1 && count <= 64 && state->n_amb == 0 && !always_rescore && !given) {  // survivors that are one computation
        __shared__ int s_dup;
        if (tid == 0) s_dup = duplicate_winner(st, sb.list, count);
        __syncthreads();
        if (s_dup >= 0) { final_one = true; cl_one = s_dup; }
    }
    if (final_one) {
        const int cl = cl_one;
        const int index = dense_of(st, given != 0, cl);
        const int li = list_index_of(st, nskip, given != 0, index, s_scratch);
        if (tid == 0) write_best(best_out, (int)(st.ws_meta[cl] & 63), st.ws_cost[cl], index, state->n_scored, li);
    }
    if (tid == 0) {
        sb.hdr[0] = final_one ? 0 : count;
        sb.hdr[1] = state->n_scored;
        sb.hdr[2] = 0;
        sb.counters[0] = 0;
        sb.counters[1] = 0;
        if (tile_counter) *tile_counter = 0;
    }
}

}  // namespace etp
