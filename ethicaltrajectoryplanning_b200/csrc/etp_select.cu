// Selection of the best candidate (sm_100a): highest non-empty validity level, then cost, then generation order
// (frenet_functions.py:562-593, frenet_planner.py:457,555-558), exact although the fused kernel scores in fp32.
//
//   select_kernel  : scans the per-candidate workspace the fused kernel leaves behind (total cost, a bound on what fp32
//                    rounding can have moved it by, level, "a decision of this candidate depends on rounding").  In fp64
//                    mode (bound 0) the minimum IS the selection.  In fp32 mode it collects every candidate that can
//                    still be the winner: those whose decisions are ambiguous, and those of the top level whose lower
//                    bound cost - bound does not exceed the smallest upper bound cost + bound.  One survivor, nothing
//                    ambiguous, no other shard: it is the winner.  Otherwise:
//   rescore_kernel : one CTA per surviving candidate re-scores it in fp64 from the exact inputs with the reference's own
//                    formulas (atan2 impact angles, Genz' BVU in MVNDST's corner order), threads over (sample, obstacle)
//                    pairs and over the 36 BVU evaluations of an un-gated pair; level and cost are re-decided, the
//                    materialised rows of the candidate are overwritten, and the last CTA takes the minimum of the exact
//                    keys.  The same idea as gate_exact / band_coefs_exact of the fused kernel, applied to the arg-min
//                    and to the max-risk / maximin-gate decisions.
#ifdef ETP_PHASE_CLOCK
#include <cstdio>
#endif
#include "etp_select.cuh"

namespace etp {

__global__ void __launch_bounds__(kSelThreads) select_kernel(const __grid_constant__ DevStep st, const SelectBufs sb, BestRec* __restrict__ best_out,
                                                            const int* __restrict__ nskip, unsigned* __restrict__ tile_counter, int given,
                                                            int always_rescore) {
    select_body(st, sb, best_out, nskip, tile_counter, given, always_rescore, blockIdx.x, gridDim.x);
}

__global__ void __launch_bounds__(kSelThreads) select_collect_kernel(const __grid_constant__ DevStep st, const SelectBufs sb,
                                                                    BestRec* __restrict__ best_out, const int* __restrict__ nskip,
                                                                    unsigned* __restrict__ tile_counter, int given, int always_rescore) {
    select_collect_body(st, sb, best_out, nskip, tile_counter, given, always_rescore, blockIdx.x, gridDim.x);
}

// scenario sweep: one block per scenario (at most kSelChunk candidates each)
__global__ void __launch_bounds__(kSelThreads) select_batch_kernel(const DevStep* __restrict__ steps, const SelectBufs* __restrict__ sbs,
                                                                  BestRec* __restrict__ bests, unsigned* __restrict__ tile_counter) {
    __shared__ DevStep s_st;
    __shared__ SelectBufs s_sb;
    const int scn = blockIdx.x;
    for (int i = threadIdx.x; i < (int)(sizeof(DevStep) / sizeof(int)); i += blockDim.x)
        reinterpret_cast<int*>(&s_st)[i] = reinterpret_cast<const int*>(steps + scn)[i];
    if (threadIdx.x == 0) s_sb = sbs[scn];
    __syncthreads();
    select_body(s_st, s_sb, bests + scn, s_st.nskip, scn == 0 ? tile_counter : nullptr, 0, 0, 0, 1);
}

// ---------------------------------------------------------------------------------------------
// rescore_kernel
// ---------------------------------------------------------------------------------------------
struct ObsSample {
    double mx, my, psi, vo, ke, ko;
    bool prot;
};

__device__ __forceinline__ ObsSample load_obs(const DevStep& st, int o, int t) {
    ObsSample s;
    const double* rp = st.raw_pos + ((size_t)o * st.Tp + t) * 2;
    s.mx = rp[0]; s.my = rp[1];
    s.psi = st.raw_yaw[(size_t)o * st.Tp + t];
    s.vo = st.raw_vel[(size_t)o * st.Tp + t];
    s.ke = st.obs_const[o * OC_COUNT + OC_KE];
    s.ko = st.obs_const[o * OC_COUNT + OC_KO];
    s.prot = st.obs_const[o * OC_COUNT + OC_PROT] != 0.0;
    return s;
}

// harm logits of (ego sample t, prediction sample t): harm_estimation.py:313-332, the reference's own expressions
__device__ __forceinline__ void harm_logits(const Tab<double>& tab, const ObsSample& s, double ex, double ey, double eyaw, double ev,
                                            double& ae, double& ao) {
    const double dv = delta_v(0.0, 0.0, ev, eyaw, s.vo, 0.0, 0.0, s.psi);
    if (s.prot) {
        double ce, co;
        bool amb = false;
        impact_coefs(tab, ex, ey, 0.0, 0.0, eyaw, s.mx, s.my, 0.0, 0.0, s.psi, ce, co, amb);
        ae = tab.lr_speed * (s.ke * dv) + ce;
        ao = tab.lr_speed * (s.ko * dv) + co;
    } else {
        ae = tab.lr1s_speed * (s.ke * dv);
        ao = tab.ped_speed * (s.ko * dv);
    }
}

// CTA `first` of `nblk` takes every nblk-th survivor of the step `st`
// Returns true in the CTA that closed the selection (it wrote *best_out).
template <bool kGiven>
__device__ __forceinline__ bool rescore_body(const DevStep& st, const DevParams& pr, const DevOut& out, int out_f32, const SelectBufs& sb,
                                             BestRec* __restrict__ best_out, const int* __restrict__ nskip, unsigned char* smem_raw,
                                             const int first, const int nblk, const int qcap = kQChunk) {
    using U = unsigned long long;
    const int count = sb.hdr[0];
    if (count == 0) return false;  // the selection was final
    const int Ts = st.Tmax;
    const int O = st.has_pred ? st.O : 0;
    const int On = O > 0 ? O : 1;
    const int tid = threadIdx.x;
    double* e_x = reinterpret_cast<double*>(smem_raw);
    double* e_y = e_x + Ts;
    double* e_yaw = e_y + Ts;
    double* e_v = e_yaw + Ts;
    double* e_c = e_v + Ts;
    double* e_s = e_c + Ts;
    U* mx = reinterpret_cast<U*>(e_s + Ts);                       // [4][On]: ego risk, obstacle risk, ego logit, obstacle logit
    double* buf = reinterpret_cast<double*>(mx + 4 * On);        // [kQChunk][36] BVU values
    double* nodes = buf + qcap * 36;                          // [kQChunk][40] sin(theta_i) | 1 - sin^2(theta_i)
    double* lims = nodes + qcap * 40;                         // [kQChunk][72] 36 rectangle limits | their normal tails
    double* ent_c = lims + qcap * 72;                         // [kQChunk][kEntC] 1/sx, 1/sy, rho, asin(rho), ego harm, obstacle harm
    unsigned* queue = reinterpret_cast<unsigned*>(ent_c + qcap * kEntC);  // [(Ts-1) * On]
    __shared__ Tab<double> tab;
    __shared__ int s_qn, s_last;
    __shared__ double s_scratch[kResThreads / 32];
    __shared__ double s_sum5[5 * (kResThreads / 32)];
    if (tid == 0) fill_tab<double>(tab, pr, 0.0);
    const bool mahal = pr.mahalanobis != 0;
    RescoreRec* res = reinterpret_cast<RescoreRec*>(sb.res);

#ifdef ETP_PHASE_CLOCK
    long long rclk[6] = {0, 0, 0, 0, 0, 0};
#define ETP_RCLK(k) do { if (tid == 0 && first == 0 && rclk[k] == 0) rclk[k] = clock64(); } while (0)
#else
#define ETP_RCLK(k) do { } while (0)
#endif
    for (int i = first; i < count; i += nblk) {
        __syncthreads();  // previous candidate consumed; tab visible
        ETP_RCLK(0);
        const int cl = sb.list[i];
        const unsigned char meta = st.ws_meta[cl];
        const int kl = kGiven ? 0 : cl / st.nd, id = kGiven ? 0 : cl - kl * st.nd;
        const int pl = kGiven ? cl : st.shard_index + kl * st.shard_count;
        const int c_dense = kGiven ? cl : (st.p_begin + pl) * st.nd + id;
        const double* gp = kGiven ? st.given + cl : st.prof + (size_t)pl * PF_COUNT * Ts;
        const double* pm = st.prof_meta + (size_t)pl * PM_COUNT;
        const bool low = kGiven ? false : pm[PM_LOW] != 0.0;
        const int T = kGiven ? st.given_T[cl] : (int)pm[PM_T];
        if (tid == 0) s_qn = 0;
        for (int k = tid; k < 4 * On; k += kResThreads) mx[k] = Enc<double>::enc(-INFINITY);

        // ---- the candidate's samples and the sums of the cost terms (frenet_functions.py:336-435) ----
        double sv = 0, sd = 0, jl = 0, jq = 0;
        {
            Poly5 q{};
            if (!kGiven) q = lateral_poly(st, low, st.d_list[id], pm[PM_TEND], pm[PM_DS]);
            for (int t = tid; t < T; t += kResThreads) {
                const TrajPoint tp = kGiven ? given_point(gp, Ts, st.given_stride, t) : traj_point(q, low, gp, Ts, t, st.dt);
                e_x[t] = tp.f[ETP_F_X]; e_y[t] = tp.f[ETP_F_Y]; e_yaw[t] = tp.f[ETP_F_YAW]; e_v[t] = tp.f[ETP_F_V];
                e_c[t] = tp.cos_yaw; e_s[t] = tp.sin_yaw;
                sv += tp.f[ETP_F_V];
                sd += fabs(tp.f[ETP_F_D]);
                jl += tp.f[ETP_F_S_DDD] * tp.f[ETP_F_S_DDD];
                jq += tp.f[ETP_F_D_DDD] * tp.f[ETP_F_D_DDD];
            }
        }
        CandFacts f;
        __syncthreads();  // every sample is in shared memory (the travelled distance reads a neighbour's)
        double trav = 0;
        for (int t = 1 + tid; t < T; t += kResThreads) {  // calc_trajectory_cost.py:739-757
            const double dx = e_x[t - 1] - e_x[t], dy = e_y[t - 1] - e_y[t];
            trav += sqrt(dx * dx + dy * dy);
        }
        {
            double sums[5] = {sv, sd, jl, jq, trav};
            block_sum_n<5>(sums, s_sum5);  // the fixed tree of block_sum, one pair of barriers for the five sums
            f.sv = sums[0]; f.sd = sums[1]; f.jl = sums[2]; f.jq = sums[3]; f.trav = sums[4];
        }
        f.T = T;
        f.x_end = e_x[T - 1];
        f.y_end = e_y[T - 1];

        ETP_RCLK(1);
        // ---- harm logits of every (t, o) sample; probability samples inside the 5 m gate are queued ----
        for (int it = tid; it < (Ts - 1) * O; it += kResThreads) {
            const int t = it / O, o = it - t * O;
            const int n_pred = st.pred_len[o];
            if (t < T - 1 && t < n_pred) {  // harm_estimation.py:234
                const ObsSample s = load_obs(st, o, t);
                double ae, ao;
                harm_logits(tab, s, e_x[t], e_y[t], e_yaw[t], e_v[t], ae, ao);
                atomicMax(mx + 2 * On + o, Enc<double>::enc(ae));
                atomicMax(mx + 3 * On + o, Enc<double>::enc(ao));
            }
            if (t < T - 1 && t + 1 < n_pred) {  // collision_probability.py:194-203: ego[t+1], mean[t], orientation[t+1]
                bool ungated = true;
                if (!mahal) {
                    const double* rp = st.raw_pos + ((size_t)o * st.Tp + t) * 2;
                    const double y1 = st.raw_yaw[(size_t)o * st.Tp + t + 1];
                    const double len = 2 * st.obs_const[o * OC_COUNT + OC_HALFLEN];
                    const double fx = cos(y1) * len / 2, fy = sin(y1) * len / 2;
                    const double xe = e_x[t + 1], ye = e_y[t + 1];
                    double ax = rp[0] - xe, ay = rp[1] - ye;
                    double m = sqrt(ax * ax + ay * ay);
                    ax = (rp[0] + fx) - xe; ay = (rp[1] + fy) - ye;
                    m = fmin(m, sqrt(ax * ax + ay * ay));
                    ax = (rp[0] - fx) - xe; ay = (rp[1] - fy) - ye;
                    m = fmin(m, sqrt(ax * ax + ay * ay));
                    ungated = !(m > 5.0);
                }
                if (ungated) queue[atomicAdd(&s_qn, 1)] = ((unsigned)t << 16) | (unsigned)o;
            }
        }
        __syncthreads();

        ETP_RCLK(2);
        // ---- queued samples: nine rectangles x four BVU corners each (collision_probability.py:205-252) ----
        const int qn = s_qn;
        for (int q0 = 0; q0 < qn; q0 += qcap) {  // qcap <= kQChunk pairs at a time (the sweep's kernel keeps its footprint small)
            const int qc = min(qcap, qn - q0);
            if (!mahal) {
                // What the 36 BVU evaluations of an entry share is computed once per entry: the Gauss-Legendre nodes of the
                // correlation (sin and 1 - sin^2, Genz' loop takes 2 * lg of them), the 18 + 18 rectangle limits and their
                // normal tails.  The sums below run in bvu_d's own order, so the values are the check mode's.
                for (int q = tid; q < qc; q += kResThreads) {
                    const unsigned ent = queue[q0 + q];
                    const int t = (int)(ent >> 16), o = (int)(ent & 0xffffu);
                    const double* cv = st.raw_cov + ((size_t)o * st.Tp + t) * 4;
                    double c00 = cv[0], c01 = cv[1], c10 = cv[2], c11 = cv[3];
                    if (c00 == 0.0 && c01 == 0.0 && c10 == 0.0 && c11 == 0.0) { c00 = 0.1; c11 = 0.1; }  // :210-212
                    const double sx = sqrt(c00), sy = sqrt(c11);
                    double* e = ent_c + q * kEntC;
                    e[0] = 1.0 / sx; e[1] = 1.0 / sy; e[2] = c10 / sy / sx;
                    e[3] = asin(e[2]);
                }
            }
            // harm of the pair's sample (risk = harm x probability, risk_costs.py:88-95): one thread per entry, next to the
            // entry constants instead of after the barrier that closes the BVU values
            static_assert(2 * kQChunk <= kResThreads, "entry constants and entry harms are computed by disjoint threads");
            for (int q = tid - kResThreads / 2; q >= 0 && q < qc; q += kResThreads) {  // the upper half of the block: runs next to the constants
                const unsigned ent = queue[q0 + q];
                const int t = (int)(ent >> 16), o = (int)(ent & 0xffffu);
                const ObsSample s = load_obs(st, o, t);
                double ae, ao;
                harm_logits(tab, s, e_x[t], e_y[t], e_yaw[t], e_v[t], ae, ao);
                double* e = ent_c + q * kEntC;
                e[4] = sigmoid_harm<double>(s.prot ? tab.lr_const : tab.lr1s_const, ae);
                e[5] = sigmoid_harm<double>(s.prot ? tab.lr_const : tab.neg_ped_const, ao);
            }
            if (!mahal) {
                __syncthreads();
                for (int item = tid; item < qc * 56; item += kResThreads) {
                    const int q = item / 56, m = item - q * 56;
                    const unsigned ent = queue[q0 + q];
                    const int t = (int)(ent >> 16), o = (int)(ent & 0xffffu);
                    const double* e = ent_c + q * kEntC;
                    if (m < 20) {  // node m of the rho integral (only the 2 * lg first are used)
                        const double ar = fabs(e[2]);
                        const int ng = ar < 0.3 ? 0 : (ar < 0.75 ? 1 : 2), i = m >> 1;
                        const double sn = sin(e[3] * (((m & 1) ? kGLx[ng][i] : -kGLx[ng][i]) + 1.0) / 2.0);
                        nodes[q * 40 + m] = sn;
                        nodes[q * 40 + 20 + m] = 1.0 - sn * sn;
                    } else {       // limit m - 20: x limits of the nine rectangles (lower, upper), then the y limits
                        const int l = m - 20, axis = l / 18, r = (l - axis * 18) >> 1, up = l & 1;
                        const int km = r / 3, jb = r - km * 3;
                        const double* rp = st.raw_pos + ((size_t)o * st.Tp + t) * 2;
                        const double y1 = st.raw_yaw[(size_t)o * st.Tp + t + 1];
                        const double len = 2 * st.obs_const[o * OC_COUNT + OC_HALFLEN];
                        const double sg = km == 0 ? 0.0 : (km == 1 ? 1.0 : -1.0), sbx = jb == 0 ? 0.0 : (jb == 1 ? 1.0 : -1.0);
                        double lim;
                        if (axis == 0) {
                            const double exo = cos(y1) * len / 2, ddx = e_x[t + 1] - rp[0], ox = tab.rr * e_c[t + 1];
                            const double bx = (ddx - sg * exo) + sbx * ox;
                            lim = (up ? bx + tab.hx : bx - tab.hx) * e[0];
                        } else {
                            const double eyo = sin(y1) * len / 2, ddy = e_y[t + 1] - rp[1], oy = tab.rr * e_s[t + 1];
                            const double by = (ddy - sg * eyo) + sbx * oy;
                            lim = (up ? by + tab.hy : by - tab.hy) * e[1];
                        }
                        lims[q * 72 + l] = lim;
                        lims[q * 72 + 36 + l] = phi_d(-lim);
                    }
                }
                __syncthreads();
                for (int item = tid; item < qc * 36; item += kResThreads) {
                    const int q = item / 36, k = item - q * 36;
                    const double* e = ent_c + q * kEntC;
                    const int r = k >> 2, cn = k & 3;
                    const int ih = 2 * r + (cn & 1), ik = 18 + 2 * r + ((cn >> 1) & 1);
                    const double h = lims[q * 72 + ih], kk = lims[q * 72 + ik], rho = e[2];
                    const double ar = fabs(rho);
                    double val;
                    if (ar < 0.925) {  // bvu_d's first branch with the shared nodes
                        const int ng = ar < 0.3 ? 0 : (ar < 0.75 ? 1 : 2), lg = ng == 0 ? 3 : (ng == 1 ? 6 : 10);
                        const double hk = h * kk, hs = (h * h + kk * kk) / 2.0;
                        const double* sn = nodes + q * 40;
                        double bvn = 0.0;
                        for (int i = 0; i < lg; ++i) {
                            bvn += kGLw[ng][i] * exp((sn[2 * i] * hk - hs) / sn[20 + 2 * i]);
                            bvn += kGLw[ng][i] * exp((sn[2 * i + 1] * hk - hs) / sn[20 + 2 * i + 1]);
                        }
                        val = bvn * e[3] / (2.0 * 6.283185307179586) + lims[q * 72 + 36 + ih] * lims[q * 72 + 36 + ik];
                    } else {
                        val = bvu_d(h, kk, rho);
                    }
                    buf[q * 36 + k] = val;
                }
                __syncthreads();
            }
            if (mahal) __syncthreads();  // the entries' harm values
            if (tid < qc) {
                const unsigned ent = queue[q0 + tid];
                const int t = (int)(ent >> 16), o = (int)(ent & 0xffffu);
                double prob = 0.0;
                if (mahal) {
                    const double* rp = st.raw_pos + ((size_t)o * st.Tp + t) * 2;
                    const double* cv = st.raw_cov + ((size_t)o * st.Tp + t) * 4;
                    const double det = cv[0] * cv[3] - cv[1] * cv[2];
                    prob = inv_mahalanobis<double>(e_x[t + 1], e_y[t + 1], rp[0], rp[1], cv[3] / det, -(cv[1] + cv[2]) / det, cv[0] / det);
                } else {
                    const double* b = buf + tid * 36;
                    for (int r = 0; r < 9; ++r) prob += b[4 * r] - b[4 * r + 1] - b[4 * r + 2] + b[4 * r + 3];  // MVNDST's corner order
                    prob = prob / 3;
                }
                const double he = ent_c[tid * kEntC + 4], ho = ent_c[tid * kEntC + 5];
                atomicMax(mx + o, Enc<double>::enc(he * prob));  // risk_costs.py:88-95
                atomicMax(mx + On + o, Enc<double>::enc(ho * prob));
            }
            __syncthreads();
        }

        ETP_RCLK(3);
        // ---- per-obstacle maxima -> principles, level, cost: the harm sigmoids by one thread per obstacle, the sums by
        // thread 0 with the obstacles in order ----------
        double* ob = buf;  // [4][On] ego risk, obstacle risk, ego harm, obstacle harm (the BVU buffer is free again; 4 On <= 36 qcap by the launcher)
        for (int o = tid; o < O; o += kResThreads) {
            const bool prt = st.obs_const[o * OC_COUNT + OC_PROT] != 0.0;
            const double re = fmax(Enc<double>::dec(mx[o]), 0.0), ro = fmax(Enc<double>::dec(mx[On + o]), 0.0);
            const double he = sigmoid_harm<double>(prt ? tab.lr_const : tab.lr1s_const, Enc<double>::dec(mx[2 * On + o]));
            const double ho = sigmoid_harm<double>(prt ? tab.lr_const : tab.neg_ped_const, Enc<double>::dec(mx[3 * On + o]));
            ob[o] = re; ob[On + o] = ro; ob[2 * On + o] = he; ob[3 * On + o] = ho;
            if (out.red) {
                const size_t Cp = out.c_stride;
                const double v[4] = {re, ro, he, ho};
                for (int k = 0; k < 4; ++k) {
                    if (out_f32) reinterpret_cast<float*>(out.red)[((size_t)k * O + o) * Cp + cl] = (float)v[k];
                    else reinterpret_cast<double*>(out.red)[((size_t)k * O + o) * Cp + cl] = v[k];
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            RiskSums r{0, 0, 0, 0, -INFINITY, -INFINITY};
            const bool gate = (pr.relaxed & ETP_RELAX_Q3_MAXIMIN) != 0;
            for (int o = 0; o < O; ++o) {
                const double re = ob[o], ro = ob[On + o], he = ob[2 * On + o], ho = ob[3 * On + o];
                r.sum_e += re;
                r.sum_o += ro;
                r.sum_abs += fabs(re - ro);
                r.resp -= st.obs_const[o * OC_COUNT + OC_RESP] * ro;
                r.mm = fmax(r.mm, he * (((re < 10e-10) != gate) ? 1.0 : 0.0));
                r.mm = fmax(r.mm, ho * (((ro < 10e-10) != gate) ? 1.0 : 0.0));
                r.max_o = fmax(r.max_o, ro);
            }
            const int level32 = meta & 63;
            f.pre_level = level32 <= 2 ? level32 : -1;  // the tests before the risk are exact in the fused kernel
            f.pre_reason = ETP_REASON_NONE;
            f.bd = st.ws_bd[cl];
            f.gf = st.ws_gf[cl];
            f.factor = st.factor_list ? st.factor_list[c_dense] : st.factor;
            double c[ETP_NUM_COST];
            int level, reason;
            assemble_costs(st, pr, O, f, r, c, level, reason);
            if (out.cost) {
                const size_t Cp = out.c_stride;
                for (int k = 0; k < ETP_NUM_COST; ++k) {
                    if (out_f32) reinterpret_cast<float*>(out.cost)[(size_t)k * Cp + cl] = (float)c[k];
                    else reinterpret_cast<double*>(out.cost)[(size_t)k * Cp + cl] = c[k];
                }
            }
            if (f.pre_level < 0) {
                if (out.level) out.level[cl] = (uint8_t)level;
                if (out.reason) out.reason[cl] = (uint8_t)reason;
            }
            // how the fp32 pass did against its own bound (diagnostics: etp_debug_counters)
            if (!(meta & SEL_AMB) && level == level32) {
                const double err = fabs(c[ETP_C_TOTAL] - st.ws_cost[cl]), b = (double)st.ws_bound[cl];
                if (err > b) atomicAdd(&sb.stats[1], 1ull);
                if (b > 0.0) atomicMax(&sb.stats[2], (unsigned long long)__float_as_uint((float)(err / b)));
            }
            atomicAdd(&sb.stats[0], 1ull);
            RescoreRec rr;
            rr.cost = c[ETP_C_TOTAL]; rr.rank = level_rank(level); rr.level = level; rr.cl = cl; rr.pad = 0;
            res[i] = rr;
        }
        ETP_RCLK(4);
#ifdef ETP_PHASE_CLOCK
        if (tid == 0 && first == 0 && i == 0)
            printf("rescore CTA 0, first survivor: %d queued pairs; cycles samples+sums %lld, harm logits + gate %lld, probabilities %lld, assembly %lld\n",
                   s_qn, rclk[1] - rclk[0], rclk[2] - rclk[1], rclk[3] - rclk[2], rclk[4] - rclk[3]);
#endif
    }

    // ---- the last CTA takes the minimum of the exact keys ----------------------------------------
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        s_last = atomicAdd(&sb.counters[1], 1u) == (unsigned)(nblk - 1);
    }
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    __shared__ int s_r[kResThreads / 32], s_c[kResThreads / 32];
    __shared__ double s_k[kResThreads / 32];
    int br = -1, bc = 0x7fffffff;
    double bk = INFINITY;
    for (int i = tid; i < count; i += kResThreads) {
        const volatile RescoreRec* r = res + i;
        const int rk = r->rank, c = r->cl;
        const double k = order_cost(r->cost);
        if (rk > br || (rk == br && (k < bk || (k == bk && c < bc)))) { br = rk; bk = k; bc = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const int orr = __shfl_xor_sync(kFull, br, o), oc = __shfl_xor_sync(kFull, bc, o);
        const double ok = __shfl_xor_sync(kFull, bk, o);
        if (orr > br || (orr == br && (ok < bk || (ok == bk && oc < bc)))) { br = orr; bk = ok; bc = oc; }
    }
    if ((tid & 31) == 0) { s_r[tid >> 5] = br; s_c[tid >> 5] = bc; s_k[tid >> 5] = bk; }
    __syncthreads();
    for (int w = 0; w < kResThreads / 32; ++w)
        if (s_r[w] > br || (s_r[w] == br && (s_k[w] < bk || (s_k[w] == bk && s_c[w] < bc)))) { br = s_r[w]; bk = s_k[w]; bc = s_c[w]; }
    // the winner's own record (cost as computed, level)
    __syncthreads();
    for (int i = tid; i < count; i += kResThreads) {
        const volatile RescoreRec* r = res + i;
        if (r->cl == bc) { s_k[0] = r->cost; s_r[0] = r->level; }
    }
    __syncthreads();
    const double wcost = s_k[0];
    const int wlevel = s_r[0];
    const int index = dense_of(st, kGiven, bc);
    const int li = list_index_of(st, nskip, kGiven, index, s_scratch);
    if (tid == 0) {
        write_best(best_out, wlevel, wcost, index, sb.hdr[1], li);
        sb.hdr[0] = 0;
        sb.counters[1] = 0;
    }
    return true;
}

template <bool kGiven>
__global__ void __launch_bounds__(kResThreads) rescore_kernel(const __grid_constant__ DevStep st, const __grid_constant__ DevParams pr,
                                                             const DevOut out, int out_f32, const SelectBufs sb, BestRec* __restrict__ best_out,
                                                             const int* __restrict__ nskip, int qcap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    rescore_body<kGiven>(st, pr, out, out_f32, sb, best_out, nskip, smem_raw, blockIdx.x, gridDim.x, qcap);
}

// scenario sweep: gridDim.y CTAs per scenario share that scenario's survivors (usually none: they leave at once)
__global__ void __launch_bounds__(kResThreads) rescore_batch_kernel(const DevStep* __restrict__ steps, const __grid_constant__ DevParams pr,
                                                                   const SelectBufs* __restrict__ sbs, BestRec* __restrict__ bests) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ DevStep s_st;
    __shared__ SelectBufs s_sb;
    const int scn = blockIdx.x;
    if (sbs[scn].hdr[0] == 0) return;  // the selection of this scenario was final
    for (int i = threadIdx.x; i < (int)(sizeof(DevStep) / sizeof(int)); i += blockDim.x)
        reinterpret_cast<int*>(&s_st)[i] = reinterpret_cast<const int*>(steps + scn)[i];
    if (threadIdx.x == 0) s_sb = sbs[scn];
    __syncthreads();
    rescore_body<false>(s_st, pr, DevOut{}, 1, s_sb, bests + scn, s_st.nskip, smem_raw, blockIdx.y, gridDim.y, kQChunkBatch);
}

// Planning cycle (etp_plan_cycle): what follows the selection in ONE launch.  The survivors of the fp32 selection are
// re-scored in fp64 by the grid's CTAs (one survivor each, round-robin); the CTA that closes the selection -- CTA 0 when
// nothing had to be re-scored -- generates the winner's trajectory and writes [BestRec | T | 13 rows] to `combo`, which may
// be pinned host memory (the cycle's only result: no separate read-back).
__global__ void __launch_bounds__(kResThreads) cycle_tail_kernel(const DevStep* __restrict__ steps, const __grid_constant__ DevParams pr,
                                                                const SelectBufs* __restrict__ sbs, BestRec* __restrict__ bests,
                                                                unsigned char* __restrict__ combo, int qcap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ DevStep s_st;
    __shared__ SelectBufs s_sb;
    __shared__ BestRec s_best;
    for (int i = threadIdx.x; i < (int)(sizeof(DevStep) / sizeof(int)); i += blockDim.x)
        reinterpret_cast<int*>(&s_st)[i] = reinterpret_cast<const int*>(steps)[i];
    if (threadIdx.x == 0) s_sb = sbs[0];
    __syncthreads();
    bool closing;
    if (s_sb.hdr[0] == 0) closing = blockIdx.x == 0;  // uniform over the grid: every CTA reads the count before any can reset it
    else closing = rescore_body<false>(s_st, pr, DevOut{}, 1, s_sb, bests, s_st.nskip, smem_raw, blockIdx.x, gridDim.x, qcap);
    if (!closing) return;
    __syncthreads();  // the record thread 0 wrote
    if (threadIdx.x == 0) s_best = *bests;
    __syncthreads();
    best_trajectory_body(s_st, s_best, combo);
}

// ---------------------------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------------------------
size_t select_blocks_bytes(int n_local) { return sizeof(SelBlock) * (size_t)((n_local + kSelChunk - 1) / kSelChunk + 1); }
size_t rescore_rec_bytes(int n_local) { return sizeof(RescoreRec) * (size_t)(n_local > 0 ? n_local : 1); }

cudaError_t launch_select(const DevStep& st, const SelectBufs& sb, BestRec* best_out, const int* nskip, unsigned* tile_counter,
                          bool given, bool always_rescore, cudaStream_t stream) {
    const int nblk = (st.n_local + kSelChunk - 1) / kSelChunk;
    select_kernel<<<nblk > 0 ? nblk : 1, kSelThreads, 0, stream>>>(st, sb, best_out, nskip, tile_counter, given ? 1 : 0, always_rescore ? 1 : 0);
    if (nblk > 1 && st.refine)
        select_collect_kernel<<<nblk, kSelThreads, 0, stream>>>(st, sb, best_out, nskip, tile_counter, given ? 1 : 0, always_rescore ? 1 : 0);
    return cudaGetLastError();
}

static size_t rescore_smem_for(int Tmax, int O, int qcap) {
    const size_t On = O > 0 ? O : 1;
    if ((size_t)qcap * 36 < 4 * On) qcap = (int)((4 * On + 35) / 36);  // the assembly's per-obstacle values reuse the BVU buffer
    return sizeof(double) * 6 * Tmax + sizeof(unsigned long long) * 4 * On + sizeof(double) * qcap * (36 + 40 + 72 + kEntC) +
           sizeof(unsigned) * (size_t)(Tmax - 1) * On + 16;
}

// pairs per chunk of a re-scoring CTA: as many as the shared memory a CTA may have allows (a long horizon with many obstacles
// needs most of it for the pair queue)
static int rescore_chunk(int Tmax, int O) {
    int q = kQChunk;
    while (q > kQChunkBatch && rescore_smem_for(Tmax, O, q) > (size_t)200 * 1024) q >>= 1;
    return q;
}

size_t rescore_smem_bytes(int Tmax, int O) { return rescore_smem_for(Tmax, O, rescore_chunk(Tmax, O)); }

cudaError_t launch_rescore(int precision, const DevStep& st, const DevParams& pr, const DevOut& out, const SelectBufs& sb,
                           BestRec* best_out, const int* nskip, int num_sms, cudaStream_t stream) {
    const int qcap = rescore_chunk(st.Tmax, st.has_pred ? st.O : 0);
    const size_t smem = rescore_smem_for(st.Tmax, st.has_pred ? st.O : 0, qcap);
    long long grid = 2ll * num_sms;
    if (grid > st.n_local) grid = st.n_local;
    if (grid < 1) grid = 1;
    const int f32 = precision == ETP_PRECISION_FP32 ? 1 : 0;
    cudaError_t e;
    if (st.given) {
        e = cudaFuncSetAttribute(rescore_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        rescore_kernel<true><<<(int)grid, kResThreads, smem, stream>>>(st, pr, out, f32, sb, best_out, nskip, qcap);
    } else {
        e = cudaFuncSetAttribute(rescore_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        rescore_kernel<false><<<(int)grid, kResThreads, smem, stream>>>(st, pr, out, f32, sb, best_out, nskip, qcap);
    }
    return cudaGetLastError();
}

cudaError_t launch_select_batch(const DevStep* steps, const SelectBufs* sbs, BestRec* bests, unsigned* tile_counter, int n_scn,
                                cudaStream_t stream) {
    if (n_scn <= 0) return cudaSuccess;
    select_batch_kernel<<<n_scn, kSelThreads, 0, stream>>>(steps, sbs, bests, tile_counter);
    return cudaGetLastError();
}

cudaError_t launch_rescore_batch(const DevStep* steps, const DevParams& pr, const SelectBufs* sbs, BestRec* bests, int n_scn, int max_T,
                                 int max_O, cudaStream_t stream) {
    if (n_scn <= 0) return cudaSuccess;
    const size_t smem = rescore_smem_for(max_T, max_O, kQChunkBatch);
    static size_t attr = 0;
    if (smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(rescore_batch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr = smem;
    }
    rescore_batch_kernel<<<dim3(n_scn, 4), kResThreads, smem, stream>>>(steps, pr, sbs, bests);
    return cudaGetLastError();
}

cudaError_t launch_cycle_tail(const DevStep* steps, const DevParams& pr, const SelectBufs* sbs, BestRec* bests, unsigned char* combo,
                              bool refine, int Tmax, int O, cudaStream_t stream) {
    const int qcap = rescore_chunk(Tmax, O);
    const size_t smem = refine ? rescore_smem_for(Tmax, O, qcap) : 0;
    static size_t attr = 0;
    if (smem > attr) {
        cudaError_t e = cudaFuncSetAttribute(cycle_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr = smem;
    }
    cycle_tail_kernel<<<refine ? 8 : 1, kResThreads, smem, stream>>>(steps, pr, sbs, bests, combo, qcap);
    return cudaGetLastError();
}

int select_chunk() { return kSelChunk; }

}  // namespace etp
