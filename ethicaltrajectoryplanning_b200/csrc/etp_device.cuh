// Device-side structures and math shared by the scoring kernels (sm_100a).
//
// Citations are to the reference tree (TUMFTM/EthicalTrajectoryPlanning), paths relative to its root.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/etp_b200.h"

namespace etp {

// rows of the per-profile table written by profile_kernel (all fp64)
enum { PF_S = 0, PF_S_D, PF_S_DD, PF_S_DDD, PF_GX, PF_GY, PF_YAW, PF_CURV, PF_CURV_D, PF_SIN, PF_COS, PF_COUNT };
// per-profile scalars
enum { PM_DS = 0, PM_LOW, PM_T, PM_VEL_OK, PM_TEND, PM_COUNT };

// Packed obstacle tile, element type R (Vec4<R> = 16 or 32 bytes).  Time-major, one record per (t, o): all lanes
// of a warp work on the same (t, o) pair (lanes are candidates), so every tile read is one uniform address.
//   G [Tp][O][2]  gate + harm : (mx, my, ex, ey), (vo, cos psi, sin psi, wrap | psi)
//                   m = pos_list[t];  e = (length/2)(cos, sin)(orientation_list[t+1])  (collision_probability.py:175)
//                   vo = v_list[t], psi = orientation_list[t]                           (harm_estimation.py:313-328)
//                   wrap = ceil((psi - pi) / 2pi): whole turns that bring psi into (-pi, pi]  (fp32 tile; fp64 keeps psi)
//   P [Tp][O][2]  probability : (1/sx, 1/sy, rho, class), (i00, i01, i11, 0)
//                   stdevs / correlation of cov_list[t] (zero matrix -> 0.1*I, collision_probability.py:205-212);
//                   class = Gauss-Legendre node count class of |rho| (0: rho = 0, 1: 4 nodes, 2: 6 nodes, 3: fp64 path);
//                   i** = inverse of cov_list[t] for the Mahalanobis mode (collision_probability.py:279)
//   N [Tp][O][5]  float4, fp32 tile only: nodes of the rho integral s[6] | q[6] | c[6] | low parts of (mx, my)
//   C [O]         per obstacle: (m_o/(m_e+m_o), m_e/(m_e+m_o), protected, responsibility)  (harm_estimation.py:327-328)
template <typename R> struct alignas(16) Vec4 { R x, y, z, w; };
template <typename R> struct alignas(2 * sizeof(R)) Vec2 { R x, y; };

enum { RC_ZERO = 0, RC_N4 = 1, RC_N6 = 2, RC_FP64 = 3 };
constexpr int kNodeVecs = 5;

template <typename R> struct Tile {
    const Vec4<R>* G;
    const Vec4<R>* P;
    const float4* N;
    const Vec4<R>* C;
    int O;
    __host__ __device__ Tile(const void* base, int O_, int Tp) : O(O_) {
        const size_t n = (size_t)O_ * Tp;
        G = reinterpret_cast<const Vec4<R>*>(base);
        P = G + 2 * n;
        N = reinterpret_cast<const float4*>(P + 2 * n);
        C = reinterpret_cast<const Vec4<R>*>(N + kNodeVecs * n);
#ifdef ETP_TMA_STAGE
        Gom = C + O_;
#endif
    }
    static __host__ __device__ size_t bytes(int O_, int Tp) {
        const size_t n = (size_t)O_ * Tp;
        size_t b = n * (4 * sizeof(Vec4<R>) + kNodeVecs * sizeof(float4)) + (size_t)O_ * sizeof(Vec4<R>);
#ifdef ETP_TMA_STAGE
        b += 2 * n * sizeof(Vec4<R>);
#endif
        return b;
    }
#ifdef ETP_TMA_STAGE
    // build variant -DETP_TMA_STAGE: a second, OBSTACLE-major copy of the G plane, [O][Tp][2], so that the records one
    // phase-2a item reads (one obstacle, a run of samples) are one contiguous range for a bulk copy (cp.async.bulk)
    const Vec4<R>* Gom;
#endif
};

// per-obstacle constants [o][OC_COUNT], fp64 (exact inputs of the fp64 gate re-check and the principle costs)
enum { OC_KE = 0, OC_KO, OC_PROT, OC_RESP, OC_HALFLEN, OC_HALFWID, OC_MAXSTEP, OC_COUNT };  // MAXSTEP: largest |dx| + |dy| between consecutive means

struct DevParams {
    double w[ETP_NUM_WEIGHTS];
    int planner_mode, mahalanobis, multi_cost, n_thr;
    int collision_check;  // ETP_COLLISION_*, already 0 when the planner mode does not test predictions
    int relaxed;          // ETP_RELAX_* bits (0: strict reference behaviour)
    double a_max;         // vehicle_params.longitudinal.a_max (ETP_RELAX_Q2_ACCELERATION)
    double max_risk;
    double thr[ETP_MAX_ANGLE_BINS];
    double fcos[ETP_MAX_ANGLE_BINS];  // cos(thr) * |cos(thr)|: band tests on direction vectors (fp32 path)
    double coef_pos[ETP_MAX_ANGLE_BINS + 1];
    double coef_neg[ETP_MAX_ANGLE_BINS + 1];
    double lr_const, lr_speed, lr1s_const, lr1s_speed, ped_const, ped_speed;
    double veh_l, veh_w, veh_m, v_max;
    double inv_turn_radius, v_thr_curv, lat_a_max;
    // fp32 images read as constant-bank operands by the specialised scoring kernels
    float f_fcos[ETP_MAX_ANGLE_BINS], f_coef_pos[ETP_MAX_ANGLE_BINS + 1], f_coef_neg[ETP_MAX_ANGLE_BINS + 1];
    float f_lr_speed, f_lr1s_speed, f_ped_speed;
    // |yaw| up to which the un-wrapped ego impact angle (quirk Q4) bins like the geometric one: pi - last threshold
    // (negative: never; needs equal last-band coefficients on both sides)
    double wrapfree_yaw;
};

// goal context of a step (etp_goal, pre-digested on the host; calc_trajectory_cost.py:349-415, 438-519)
struct DevGoal {
    int on;                  // 0: the step's scalars / factor list are used
    int has_area, n_poly, n_circ;
    const int* poly_start;   // [n_poly + 1]
    const double* verts;     // [.][2]
    const double* circles;   // [n_circ][3]
    int has_vel, has_ori, ori_invert;
    double v_lo, v_hi, vel_mid;
    double o_lo, o_hi;       // check_orientation's normalised, sorted interval; ori_invert: exactly one end was folded
    double bx0, bx1, by0, by1; // bounding box of the goal area: points outside it are outside every shape
    int time_first;          // first sample index whose time step lies inside the goal's time window (INT_MAX: none)
    int ego_in_goal;         // reached_target_position(ego_state.position)
    int vel_kind;            // outside the goal at sample 0: 0 |vel_out - mean v|, 1 30 - mean v, 3 zero
    double vel_out;          // avg_dist / remaining_time
    // calc_dist_to_goal_pos (calc_trajectory_cost.py:760-803): centre lines of the goal lanelets or centres of the goal shapes
    int n_lines, n_pts;
    const int* line_start;   // [n_lines + 1]
    const double* line_verts;
    const double* pts;       // [n_pts][2]
};

// distance of (x, y) to the closest goal position: closest point of the goal lanelets' centre lines
// (dist_to_nearest_point, calc_trajectory_cost.py:835-853: shapely's project + interpolate, i.e. the nearest point of the
// polyline), else the closest goal-shape centre, else 0 (survival scenario)
__device__ inline double dist_to_goal_pos(const DevGoal& g, double x, double y) {
    double best = INFINITY;
    if (g.n_lines > 0) {
        for (int l = 0; l < g.n_lines; ++l) {
            const int a = g.line_start[l], b = g.line_start[l + 1];
            double dl = INFINITY;
            if (b - a == 1) {
                const double dx = x - g.line_verts[2 * a], dy = y - g.line_verts[2 * a + 1];
                dl = sqrt(dx * dx + dy * dy);
            }
            for (int i = a; i + 1 < b; ++i) {
                const double x0 = g.line_verts[2 * i], y0 = g.line_verts[2 * i + 1];
                const double ex = g.line_verts[2 * i + 2] - x0, ey = g.line_verts[2 * i + 3] - y0;
                const double l2 = ex * ex + ey * ey;
                double u = l2 > 0.0 ? ((x - x0) * ex + (y - y0) * ey) / l2 : 0.0;
                u = u < 0.0 ? 0.0 : (u > 1.0 ? 1.0 : u);
                const double dx = x - (x0 + u * ex), dy = y - (y0 + u * ey);
                dl = fmin(dl, sqrt(dx * dx + dy * dy));
            }
            best = fmin(best, dl);
        }
        return best;
    }
    if (g.n_pts > 0) {
        for (int k = 0; k < g.n_pts; ++k) {
            const double dx = g.pts[2 * k] - x, dy = g.pts[2 * k + 1] - y;
            best = fmin(best, sqrt(dx * dx + dy * dy));
        }
        return best;
    }
    return 0.0;
}

// even-odd containment of (x, y) in the union of the goal's polygons and circles.  The crossing test
// x < xi + (xj - xi) (y - yi) / (yj - yi) is evaluated without the division (multiplied through by yj - yi, the
// inequality turned for a negative one).
__host__ __device__ inline bool point_in_goal(const DevGoal& g, const int* poly_start, const double* verts, const double* circles,
                                              double x, double y) {
    if (x < g.bx0 || x > g.bx1 || y < g.by0 || y > g.by1) return false;
    for (int p = 0; p < g.n_poly; ++p) {
        const int a = poly_start[p], b = poly_start[p + 1];
        bool in = false;
        for (int i = a, j = b - 1; i < b; j = i++) {
            const double xi = verts[2 * i], yi = verts[2 * i + 1], xj = verts[2 * j], yj = verts[2 * j + 1];
            if ((yi > y) != (yj > y)) {
                const double dyj = yj - yi, lhs = (x - xi) * dyj, rhs = (xj - xi) * (y - yi);
                if (dyj > 0.0 ? lhs < rhs : lhs > rhs) in = !in;
            }
        }
        if (in) return true;
    }
    for (int k = 0; k < g.n_circ; ++k) {
        const double dx = x - circles[3 * k], dy = y - circles[3 * k + 1], r = circles[3 * k + 2];
        if (dx * dx + dy * dy <= r * r) return true;
    }
    return false;
}

// road boundary (etp_set_road_boundary): triangles in a uniform grid, fp64
struct DevBoundary {
    int check;                // ETP_COLLISION_*; HOST: nothing to test
    int n_tri, nx, ny;
    double x0, y0, inv_cell;  // cell (ix, iy) covers [x0 + ix / inv_cell, ...)
    const double* tri;        // [n_tri][6]
    const float* trif;        // [n_tri][6] relative to (x0, y0), fp32 (first pass of the fp32 kernels)
    double off_x, off_y;      // step origin - (x0, y0)
    float tol_f;              // |margin| below which the fp32 pass defers to fp64 [m]
    const int* cell_start;    // [nx * ny + 1]
    const int* cell_items;    // triangle indices per cell
    double lr1s_const, lr1s_speed;
};

struct DevStep {
    double c_s, c_s_d, c_s_dd, c_d, c_d_d, c_d_dd, dt, v_thr;
    const double* v_list;
    const double* t_list;
    const double* d_list;
    const int* nsamp;
    int nv, nt, nd, Tmax;
    int c_begin, c_end, p_begin, p_end;
    int shard_index, shard_count, n_local;  // this call scores profiles p_begin + shard_index + k * shard_count
    int vel_mode;
    double vel_target, factor;
    double eps_pos;            // fp32 rounding of a position difference in this step's coordinate range [m]
    double origin_x, origin_y; // fp32 ego samples and the obstacle tile are relative to this point
    const uint8_t* ext_level;  // device copy, indexed by dense candidate, or null
    const double* bd_harm;     // device copy, indexed by dense candidate, or null
    const double* factor_list; // device copy, indexed by dense candidate, or null (then `factor`)
    DevGoal goal;              // goal.on: velocity cost and factor derived per candidate instead
    DevBoundary boundary;      // boundary.check != HOST: level 2 / boundary harm derived per candidate instead of bd_harm
    // spline tables
    const double* knots;
    const double* cx;
    const double* cy;
    int nseg;
    // per-profile table [p - p_begin][PF_COUNT][Tmax] and scalars [p - p_begin][PM_COUNT]
    double* prof;
    double* prof_meta;
    int* nskip;              // [p - p_begin] lateral targets of the profile the reference skips (ds <= |dT|)
    // obstacles
    int has_pred;  // 0: predictions is None
    int O, Tp;
    const void* obs;         // packed obstacle tile (struct Tile<R>)
    const double* obs_const; // [O][OC_COUNT]
    const int* pred_len;     // [O]
    // raw fp64 copies for the exact gate re-check of the fp32 path
    const double* raw_pos;   // [O][Tp][2]
    const double* raw_yaw;   // [O][Tp]
    const double* raw_cov;   // [O][Tp][4]  (exact re-scoring of near-tie candidates, rescore_kernel)
    const double* raw_vel;   // [O][Tp]
    // selection workspace, one entry per local candidate column (etp_select.cu): total cost, bound on what fp32
    // rounding can have moved it by, level | SEL_AMB, boundary harm, goal flags
    double* ws_cost;
    float* ws_bound;
    uint8_t* ws_meta;
    double* ws_bd;
    uint8_t* ws_gf;
    int refine;              // 1: fp32 arithmetic, near-ties and rounding-dependent decisions are re-made in fp64
    // caller-provided trajectories instead of generated candidates (etp_score_trajectories), or null
    const double* given;     // [ETP_NUM_FIELDS][Tmax][given_stride], fp64
    const int* given_T;      // [n_local] samples per trajectory
    int given_stride;
};

// arguments of pack_obstacles_kernel: the prediction arrays of one planning step (etp_obstacles, device copies)
struct PackArgs {
    int O, Tp;
    const int* pred_len;
    const double *pos, *cov, *yaw, *vel, *length, *width, *mass;
    const int *prot, *resp;
    double veh_m, ox, oy;
    void* tile;
    double* oconst;
};

// arguments of the prediction enrichment (enrich_kernel) as a record in device memory
struct EnrichArgs {
    int O, Tp;
    const int* pred_len;
    const double *pos, *raw_len, *raw_wid, *ori0, *vel0;
    double dt, margin_l, margin_w, ego_x, ego_y, ego_ori;
    int wrap_view_cone;
    double *yaw, *vel, *length, *width;
    int* resp;
};

// scenario sweep (etp_plan_batch): the steps of a chunk and the map from the chunk's tiles to them
struct BestRec;
struct BatchArgs {
    const DevStep* steps;   // [n_steps] device copies
    const int* tile_scn;    // [n_tiles] scenario of every 32-candidate tile
    const int* tile_first;  // [n_steps] first tile of every scenario
    int n_tiles, n_steps;
    int dbg_min_len;        // experiment knob (ETP_CLUSTER_MINLEN): samples per phase-2a segment in cluster mode, 0 = default
    // planning cycle (one step, at most one selection block of candidates): the CTA that finishes last runs the selection
    const void* sel;        // SelectBufs of the step, or null: selection by its own launch
    BestRec* best;
    unsigned* done;         // completion counter of the grid (zero between launches)
};

struct DevOut {
    void* traj;
    void* prob;
    void* red;
    void* cost;
    uint8_t* level;
    uint8_t* reason;
    void* harm;
    double* bd_harm;
    size_t c_stride;  // elements between rows of the candidate-minor tensors
};

struct BestRec {
    unsigned long long key_hi, key_lo;
    double cost;
    int index, level, n_scored, list_index;
};

__host__ __device__ inline int level_rank(int level) { return level == 10 ? 4 : level; }

// monotone map of a double to u64 (smaller double -> smaller key); NaN maps to the maximum
__device__ __forceinline__ unsigned long long monotone_key(double x) {
    if (x != x) return ~0ull;
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// 128-bit selection key, smaller is better: (4 - level rank | cost | index) with the full 64-bit cost image, so that the
// key orders exactly like the reference's stable sort by cost within the highest non-empty level
// (frenet_functions.py:562-570, frenet_planner.py:457): hi = rank' << 61 | mono >> 3, lo = (mono & 7) << 61 | index
__host__ __device__ __forceinline__ unsigned long long make_key_hi_from(int level, unsigned long long mono) {
    return ((unsigned long long)(4 - level_rank(level)) << 61) | (mono >> 3);
}
__host__ __device__ __forceinline__ unsigned long long make_key_lo_from(unsigned long long mono, unsigned long long index) {
    return ((mono & 7ull) << 61) | index;
}
__device__ __forceinline__ unsigned long long make_key_hi(int level, double cost) { return make_key_hi_from(level, monotone_key(cost)); }
__device__ __forceinline__ unsigned long long make_key_lo(double cost, unsigned long long index) {
    return make_key_lo_from(monotone_key(cost), index);
}

// ---------------------------------------------------------------------------------------------
// cost assembly of one candidate (risk_costs.py:128-255, calc_trajectory_cost.py:103-148,279-346,418-519,726-757,
// validity_checks.py:277-294), shared by the fused kernel (phase 3b) and the exact re-scoring
// ---------------------------------------------------------------------------------------------
enum { GF_REACHED = 1, GF_OUTSIDE = 2, GF_START_INSIDE = 4 };  // goal bookkeeping of a candidate (goal_sample_flags)
enum { SEL_AMB = 64, SEL_SKIPPED = 255 };                       // ws_meta: level | SEL_AMB, or SEL_SKIPPED

struct RiskSums {
    double sum_e, sum_o, sum_abs, resp, mm, max_o;  // over the obstacles: ego / obstacle risk, |difference|, responsibility,
};                                                  // gated harm maximum (Q3), largest obstacle risk
struct CandFacts {
    double sv, sd, jl, jq, trav;  // sums over the samples: v, |d|, s_ddd^2, d_ddd^2, travelled distance
    double x_end, y_end;          // last position (calc_dist_to_goal_pos)
    int T;
    int pre_level, pre_reason;    // validity decided before the risk: 0 / 1 / 2 with its reason, or -1
    double bd;                    // boundary harm (risk_costs.py:104-123)
    int gf;                       // GF_* flags (st.goal.on)
    double factor;                // caller's cost factor (list entry or scalar), replaced by the goal logic if on
};

// what fp32 rounding of the risk terms can move a candidate's decisions and total cost by (DevStep::refine)
constexpr double kTolRisk = 1e-4;      // relative bound of a per-obstacle risk maximum (the parity bar of the probability)
constexpr double kTolRiskAbs = 1e-9;   // its absolute floor
constexpr double kTolLogit = 2e-5;     // absolute bound of a harm logit (fp32 velocities and headings)
constexpr double kTolSigmoid = 5e-7;   // relative bound of the fp32 sigmoid itself
constexpr double kTolGate = 1e-2;      // relative bound of a risk at the maximin gate's eps = 1e-9 (risk_costs.py:204-206): rectangles
                                       // beyond 7 sigma are dropped (< 4e-12 on the probability) and fp32 carries ~1e-4 down there

__device__ __forceinline__ void assemble_costs(const DevStep& st, const DevParams& pr, int O, const CandFacts& f, const RiskSums& r,
                                               double* c, int& level, int& reason) {
#pragma unroll
    for (int k = 0; k < ETP_NUM_COST; ++k) c[k] = 0.0;
    const double bd = f.bd;
    if (O > 0) {
        c[ETP_C_BAYES] = (r.sum_e + r.sum_o + bd) / (double)(O * 2);  // risk_costs.py:148-150
        c[ETP_C_EQUALITY] = r.sum_abs / (double)O;                    // :176-178
        {  // max(...) ** 10 (:208) as four multiplications
            const double x = fmax(r.mm, bd), x2 = x * x, x4 = x2 * x2;
            c[ETP_C_MAXIMIN] = x4 * x4 * x2;
        }
        c[ETP_C_EGO] = r.sum_e + bd;                                  // :226
        c[ETP_C_RESPONSIBILITY] = r.resp;
    }
    const bool risk_in_cost = pr.w[ETP_W_RISK_COST] > 0.0 && st.has_pred && pr.planner_mode == ETP_MODE_RISK;
    if (risk_in_cost)  // calc_trajectory_cost.py:130-136
        c[ETP_C_RISK_TOTAL] = pr.w[ETP_W_BAYES] * c[ETP_C_BAYES] + pr.w[ETP_W_EQUALITY] * c[ETP_C_EQUALITY] +
                              pr.w[ETP_W_MAXIMIN] * c[ETP_C_MAXIMIN] +
                              pr.w[ETP_W_RESPONSIBILITY] * pr.w[ETP_W_BAYES] * c[ETP_C_RESPONSIBILITY] +
                              pr.w[ETP_W_EGO] * c[ETP_C_EGO];
    const double mean_v = f.sv / (double)f.T;
    int vel_mode = st.vel_mode;
    double vel_target = st.vel_target;
    double factor = f.factor;
    if (st.goal.on) {
        const int gf = f.gf;
        if (gf & GF_START_INSIDE) {  // inside the goal area: the goal's velocity, else slow (:452-467)
            vel_mode = st.goal.has_vel ? 0 : 2;
            vel_target = st.goal.vel_mid;
        } else {                     // :469-519
            vel_mode = st.goal.vel_kind;
            vel_target = st.goal.vel_out;
        }
        factor = 1.0;                // get_cost_factor, :375-415
        bool in_time = false;
        if (gf & GF_REACHED) {
            factor = 0.01;
            in_time = st.goal.time_first < f.T;
            if (in_time) factor = 0.0001;
        }
        if (!in_time && st.goal.ego_in_goal && (gf & GF_OUTSIDE)) factor = 10.0;
    }
    switch (vel_mode) {  // calc_trajectory_cost.py:438-519 outcome forms
        case 0: c[ETP_C_VELOCITY] = fabs(vel_target - mean_v); break;
        case 1: c[ETP_C_VELOCITY] = 30.0 - mean_v; break;
        case 2: c[ETP_C_VELOCITY] = mean_v; break;
        default: c[ETP_C_VELOCITY] = 0.0;
    }
    c[ETP_C_DIST_TO_GLOBAL_PATH] = f.sd / (double)f.T;  // :726-736
    c[ETP_C_LON_JERK] = f.jl / (double)f.T;             // :418-435
    c[ETP_C_LAT_JERK] = f.jq / (double)f.T;
    c[ETP_C_TRAVELLED_DIST] = f.trav;
    if (pr.w[ETP_W_DIST_TO_GOAL_POS] > 0.0 && st.goal.on) c[ETP_C_DIST_TO_GOAL_POS] = dist_to_goal_pos(st.goal, f.x_end, f.y_end);
    c[ETP_C_FACTOR] = factor;
    // validity level (validity_checks.py:31-122): the tests before the risk have run (pre_level)
    if (f.pre_level >= 0) { level = f.pre_level; reason = f.pre_reason; }
    else if (pr.planner_mode == ETP_MODE_RISK && st.has_pred && O > 0 &&
             (r.sum_e > pr.max_risk || r.max_o > pr.max_risk)) { level = 3; reason = ETP_REASON_MAX_RISK; }
    else { level = 10; reason = ETP_REASON_NONE; }
    // weighted sum (calc_trajectory_cost.py:321-346); risk-only weights below level 10 if requested
    const bool risk_only = level < 10 && pr.multi_cost;
    double cost = 0.0;
    if (risk_in_cost) cost += pr.w[ETP_W_RISK_COST] * c[ETP_C_RISK_TOTAL];
    if (pr.w[ETP_W_LON_JERK] > 0.0) cost += (risk_only ? 0.0 : pr.w[ETP_W_LON_JERK]) * c[ETP_C_LON_JERK];
    if (pr.w[ETP_W_LAT_JERK] > 0.0) cost += (risk_only ? 0.0 : pr.w[ETP_W_LAT_JERK]) * c[ETP_C_LAT_JERK];
    if (pr.w[ETP_W_VELOCITY] > 0.0) cost += (risk_only ? 0.0 : pr.w[ETP_W_VELOCITY]) * c[ETP_C_VELOCITY];
    if (pr.w[ETP_W_DIST_TO_GLOBAL_PATH] > 0.0)
        cost += (risk_only ? 0.0 : pr.w[ETP_W_DIST_TO_GLOBAL_PATH]) * c[ETP_C_DIST_TO_GLOBAL_PATH];
    if (pr.w[ETP_W_TRAVELLED_DIST] > 0.0)
        cost += (risk_only ? 0.0 : pr.w[ETP_W_TRAVELLED_DIST]) * c[ETP_C_TRAVELLED_DIST];
    if (pr.w[ETP_W_DIST_TO_GOAL_POS] > 0.0)
        cost += (risk_only ? 0.0 : pr.w[ETP_W_DIST_TO_GOAL_POS]) * c[ETP_C_DIST_TO_GOAL_POS];
    c[ETP_C_TOTAL] = factor * cost;
}

// Bound on |fp32-path total cost - exact total cost| of a candidate from its own risk terms, and whether one of its
// decisions (max-risk level, maximin gate) is too close to its threshold for fp32 to make (DevStep::refine).
__device__ __forceinline__ float cost_bound(const DevParams& pr, int O, const RiskSums& r, double bd, const double* c, bool risk_in_cost) {
    if (!risk_in_cost || O <= 0) return 0.0f;
    const double x = fmax(r.mm, bd);
    const double mterm = c[ETP_C_MAXIMIN];
    const double d_risk = kTolRisk * (r.sum_e + r.sum_o) + 2.0 * O * kTolRiskAbs;  // bound of sum(d ego risk) + sum(d obstacle risk)
    const double lin = (fabs(pr.w[ETP_W_BAYES]) / (2.0 * O) + fabs(pr.w[ETP_W_EQUALITY]) / O + fabs(pr.w[ETP_W_EGO]) +
                        fabs(pr.w[ETP_W_RESPONSIBILITY] * pr.w[ETP_W_BAYES])) * d_risk;
    const double mm = fabs(pr.w[ETP_W_MAXIMIN]) * mterm * (10.0 * fmax(1.0 - x, 0.0) * kTolLogit + 10.0 * kTolSigmoid);
    return (float)(1.0000001 * fabs(c[ETP_C_FACTOR] * pr.w[ETP_W_RISK_COST]) * (lin + mm)) ;
}

// ---------------------------------------------------------------------------------------------
// lateral quintic with end velocity / acceleration zero (polynomials.py:9-55, closed form of the
// 3x3 solve) and the longitudinal quartic (polynomials.py:146-177)
// ---------------------------------------------------------------------------------------------
struct Poly5 {
    double a0, a1, a2, a3, a4, a5;
    __device__ __forceinline__ double p(double t) const {
        double t2 = t * t, t3 = t2 * t, t4 = t2 * t2, t5 = t4 * t;
        return a0 + a1 * t + a2 * t2 + a3 * t3 + a4 * t4 + a5 * t5;
    }
    __device__ __forceinline__ double d1(double t) const {
        double t2 = t * t, t3 = t2 * t, t4 = t2 * t2;
        return a1 + 2 * a2 * t + 3 * a3 * t2 + 4 * a4 * t3 + 5 * a5 * t4;
    }
    __device__ __forceinline__ double d2(double t) const {
        double t2 = t * t, t3 = t2 * t;
        return 2 * a2 + 6 * a3 * t + 12 * a4 * t2 + 20 * a5 * t3;
    }
    __device__ __forceinline__ double d3(double t) const {
        double t2 = t * t;
        return 6 * a3 + 24 * a4 * t + 60 * a5 * t2;
    }
};

__device__ __forceinline__ Poly5 make_quintic(double xs, double vxs, double axs, double xe, double T) {
    Poly5 q;
    q.a0 = xs;
    q.a1 = vxs;
    q.a2 = axs / 2.0;
    double T2 = T * T, T3 = T2 * T, T4 = T2 * T2, T5 = T4 * T;
    double b0 = xe - q.a0 - q.a1 * T - q.a2 * T2;
    double b1 = 0.0 - q.a1 - 2 * q.a2 * T;
    double b2 = 0.0 - 2 * q.a2;
    q.a3 = (10.0 * b0 - 4.0 * b1 * T + 0.5 * b2 * T2) / T3;
    q.a4 = (-15.0 * b0 + 7.0 * b1 * T - b2 * T2) / T4;
    q.a5 = (6.0 * b0 - 3.0 * b1 * T + 0.5 * b2 * T2) / T5;
    return q;
}

// lateral polynomial of candidate (profile, dT): time parametrised for high velocities, arclength
// parametrised for low velocities (frenet_functions.py:336-411)
__device__ __forceinline__ Poly5 lateral_poly(const DevStep& st, bool low, double dT, double t_end, double ds) {
    if (!low) return make_quintic(st.c_d, st.c_d_d, st.c_d_dd, dT, t_end);
    double dd_nt = 0.0, ddd_nt = 0.0;
    if (st.c_s_d != 0.0) {
        dd_nt = st.c_d_d / st.c_s_d;
        ddd_nt = (st.c_d_dd - st.c_s_dd * dd_nt) / (st.c_s_d * st.c_s_d);
    }
    return make_quintic(st.c_d, dd_nt, ddd_nt, dT, ds);
}

struct TrajPoint {
    double f[ETP_NUM_FIELDS];
    double cos_yaw, sin_yaw;
    double wrap;  // ceil((yaw - pi) / 2pi)
};

// One sample of one candidate: frenet_functions.py:357-365 / 404-411 (lateral state) and 420-435
// (Frenet -> Cartesian).  `pf` points at the profile rows, stride `ts` between rows.
__device__ __forceinline__ TrajPoint traj_point(const Poly5& q, bool low, const double* pf, int ts, int t, double dt) {
    TrajPoint r;
    const double s = pf[PF_S * ts + t], s_d = pf[PF_S_D * ts + t], s_dd = pf[PF_S_DD * ts + t];
    const double kr = pf[PF_CURV * ts + t], krd = pf[PF_CURV_D * ts + t];
    const double sr = pf[PF_SIN * ts + t], cr = pf[PF_COS * ts + t];
    double d, dp, dpp, d3, d_d_time, d_dd_time;
    if (!low) {
        const double tau = t * dt;
        d = q.p(tau);
        d_d_time = q.d1(tau);
        d_dd_time = q.d2(tau);
        d3 = q.d3(tau);
        dp = d_d_time / s_d;
        dpp = (d_dd_time - dp * s_dd) / (s_d * s_d);
    } else {
        const double tau = s - pf[PF_S * ts + 0];
        d = q.p(tau);
        dp = q.d1(tau);
        dpp = q.d2(tau);
        d3 = q.d3(tau);
        d_d_time = s_d * dp;
        d_dd_time = s_dd * dp + (s_d * s_d) * dpp;
    }
    const double om = 1.0 - kr * d;
    const double z = dp / om;
    const double dpsi = atan(z);
    const double cd = 1.0 / sqrt(1.0 + z * z);  // cos(atan z)
    const double sd = z * cd;                   // sin(atan z)
    r.f[ETP_F_D] = d;
    r.f[ETP_F_D_D] = d_d_time;
    r.f[ETP_F_D_DD] = d_dd_time;
    r.f[ETP_F_D_DDD] = d3;
    r.f[ETP_F_S] = s;
    r.f[ETP_F_S_D] = s_d;
    r.f[ETP_F_S_DD] = s_dd;
    r.f[ETP_F_S_DDD] = pf[PF_S_DDD * ts + t];
    r.f[ETP_F_X] = pf[PF_GX * ts + t] - d * sr;
    r.f[ETP_F_Y] = pf[PF_GY * ts + t] + d * cr;
    r.f[ETP_F_YAW] = dpsi + pf[PF_YAW * ts + t];
    r.f[ETP_F_V] = (s_d * om) / cd;
    r.f[ETP_F_CURV] = (((dpp + (krd * d + kr * dp) * z) * ((cd * cd) / om)) + kr) * (cd / om);
    r.cos_yaw = cd * cr - sd * sr;
    r.sin_yaw = sd * cr + cd * sr;
    r.wrap = ceil((r.f[ETP_F_YAW] - 3.14159265358979323846) / 6.28318530717958647692);
    return r;
}

// The winner of a step with its trajectory, all threads of the block: [BestRec | int T | pad | fields[ETP_NUM_FIELDS][Tmax]]
// (frenet_planner.py:564-576 reads these rows); requires the winner's profile in st.prof.
__device__ __forceinline__ void best_trajectory_body(const DevStep& st, const BestRec& b, unsigned char* __restrict__ combo) {
    BestRec* b_out = reinterpret_cast<BestRec*>(combo);
    int* n_out = reinterpret_cast<int*>(combo + sizeof(BestRec));
    double* fields = reinterpret_cast<double*>(combo + sizeof(BestRec) + 8);
    if (threadIdx.x == 0) *b_out = b;
    if (b.index < 0) {
        if (threadIdx.x == 0) *n_out = 0;
        return;
    }
    const int p = b.index / st.nd, id = b.index % st.nd;
    const int pl = p - st.p_begin;
    const double* gp = st.prof + (size_t)pl * PF_COUNT * st.Tmax;
    const double* pm = st.prof_meta + (size_t)pl * PM_COUNT;
    const bool low = pm[PM_LOW] != 0.0;
    const int T = (int)pm[PM_T];
    const Poly5 q = lateral_poly(st, low, st.d_list[id], pm[PM_TEND], pm[PM_DS]);
    for (int t = threadIdx.x; t < st.Tmax; t += blockDim.x) {
        if (t < T) {
            const TrajPoint tp = traj_point(q, low, gp, st.Tmax, t, st.dt);
            for (int f = 0; f < ETP_NUM_FIELDS; ++f) fields[f * st.Tmax + t] = tp.f[f];
        } else {
            for (int f = 0; f < ETP_NUM_FIELDS; ++f) fields[f * st.Tmax + t] = 0.0;
        }
    }
    if (threadIdx.x == 0) *n_out = T;
}

// One sample of a caller-provided trajectory (etp_score_trajectories): `g` points at field 0, sample 0 of the
// trajectory, rows are `ts` samples x `stride` trajectories.
__device__ __forceinline__ TrajPoint given_point(const double* g, int ts, int stride, int t) {
    TrajPoint r;
#pragma unroll
    for (int f = 0; f < ETP_NUM_FIELDS; ++f) r.f[f] = g[((size_t)f * ts + t) * stride];
    sincos(r.f[ETP_F_YAW], &r.sin_yaw, &r.cos_yaw);
    r.wrap = ceil((r.f[ETP_F_YAW] - 3.14159265358979323846) / 6.28318530717958647692);
    return r;
}

// curvature test of validity_checks.py:180-213 for one sample: 0 ok, 1 "lvc", 2 "hvc"
__device__ __forceinline__ int curvature_fail(const DevParams& pr, double v, double curv) {
    const double c = fabs(curv);
    if (v < pr.v_thr_curv) return (c > pr.inv_turn_radius) ? 1 : 0;
    const double cmax = pr.lat_a_max / (v * v);
    return (c > cmax) ? 2 : 0;
}

}  // namespace etp
