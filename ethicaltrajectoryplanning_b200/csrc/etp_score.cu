// Fused per-candidate scoring kernel of the B200-native candidate-trajectory path (sm_100a).
//
//   score_kernel<R> : quintic + Frenet->Cartesian + validity masks, collision probability, harm, risk maxima
//                     over time, ethical principle costs, weighted cost, validity level, block arg-min; the
//                     last block finishes the grid arg-min (frenet_functions.py:330-474, 529-593;
//                     collision_probability.py:136-256; harm_estimation.py:306-332; risk_costs.py:87-101,
//                     128-255; calc_trajectory_cost.py:103-146,279-346; validity_checks.py:31-122)
//
// Work decomposition ("lanes are candidates").  A tile is 32 consecutive candidates of the generation
// order, one per lane, owned by one CTA of 8 warps for its whole life:
//   phase 1  warps split the TIME axis: each lane generates its candidate's samples of the warp's time chunk
//            in fp64 (masks are bit-exact), streams the 13 trajectory fields out and leaves the ego samples
//            (x, y, cos yaw, sin yaw | v, yaw wraps) in shared memory;
//   phase 2a warps pull (obstacle group, time segment) items: every lane walks its own candidate against the
//            SAME (t, o) prediction sample, so each obstacle record is one uniform (broadcast) 16-byte load; the
//            harm logits are straight-line code with running maxima in registers; (t, o) samples with a lane
//            within reach of the 5 m gate go to a shared-memory queue;
//   phase 2b warps pull queue entries: three-point gate, bivariate-normal probability, risk.  Neighbouring
//            lateral targets are centimetres apart, so whole warps take this branch together, and all warps
//            of the CTA run the same few hundred instructions of code at the same time;
//   phase 3  warps split the OBSTACLE axis for the per-obstacle reductions, warp 0 finishes cost, level, key.
// All materialised tensors are candidate-minor ([...][c]): every store of a warp is one contiguous 128-byte
// line, written exactly once.
#ifdef ETP_PHASE_CLOCK
#include <cstdio>
#endif
#include <cstdlib>
#include <cuda_fp16.h>

#include <type_traits>

#include "etp_bvn.cuh"
#include "etp_device.cuh"
#include "etp_harm.cuh"
#include "etp_launch.h"
#include "etp_select.cuh"

namespace etp {

#ifndef ETP_THREADS
#define ETP_THREADS 256
#endif
constexpr int kThreads = ETP_THREADS;
constexpr int kWarps = kThreads / 32;
#ifndef ETP_MIN_BLOCKS
#define ETP_MIN_BLOCKS 2
#endif
constexpr int kMinBlocks = ETP_MIN_BLOCKS;  // resident CTAs per SM the register budget is sized for
#ifndef ETP_SEG_FACTOR
#define ETP_SEG_FACTOR 2  // phase-2a items per warp (2: 17.0 ms, 3: 17.1, 4: 17.2, 6: 17.6 on the 10^6-candidate step)
#endif
#ifndef ETP_OB
#define ETP_OB 4
#endif
constexpr int kOB = ETP_OB;  // obstacles of one phase-2 item (register blocking: one ego load serves kOB obstacles)
constexpr float kLoScale = 16384.0f;  // low parts of coordinates are stored as halves of (residue * 2^14)

// ---------------------------------------------------------------------------------------------
// exact (fp64) impact-area coefficients of one (candidate, sample, obstacle): the fp32 paths call this when a band
// test is too close to its threshold to be decided in fp32.  Everything is recomputed from the exact inputs the way
// the reference does it: rel = atan2(dy, dx), angles rel - yaw_e and pi + rel - psi_o, un-wrapped
// (harm_estimation.py:313-319; logistic_regression_*.py).
// ---------------------------------------------------------------------------------------------
struct ExactRef {
    const DevStep* st;
    const DevParams* pr;
    bool given;
    int pl, id, t, o;
};

__device__ __forceinline__ double angle_coef_exact(const DevParams& pr, double angle) {
    const double a = fabs(angle);
    int i = 0;
    while (i < pr.n_thr && !(a < pr.thr[i])) ++i;
    return angle < 0.0 ? pr.coef_neg[i] : pr.coef_pos[i];
}

__device__ __noinline__ void band_coefs_exact(const ExactRef& r, double& ce, double& co) {
    const DevStep& st = *r.st;
    double xe, ye, yaw;
    if (r.given) {
        const double* g = st.given + r.pl;
        xe = g[((size_t)ETP_F_X * st.Tmax + r.t) * st.given_stride];
        ye = g[((size_t)ETP_F_Y * st.Tmax + r.t) * st.given_stride];
        yaw = g[((size_t)ETP_F_YAW * st.Tmax + r.t) * st.given_stride];
    } else {
        const double* gp = st.prof + (size_t)r.pl * PF_COUNT * st.Tmax;
        const double* pm = st.prof_meta + (size_t)r.pl * PM_COUNT;
        const bool low = pm[PM_LOW] != 0.0;
        const Poly5 q = lateral_poly(st, low, st.d_list[r.id], pm[PM_TEND], pm[PM_DS]);
        const TrajPoint tp = traj_point(q, low, gp, st.Tmax, r.t, st.dt);
        xe = tp.f[ETP_F_X];
        ye = tp.f[ETP_F_Y];
        yaw = tp.f[ETP_F_YAW];
    }
    const double* rp = st.raw_pos + ((size_t)r.o * st.Tp + r.t) * 2;
    const double psi = st.raw_yaw[(size_t)r.o * st.Tp + r.t];
    const double rel = atan2(rp[1] - ye, rp[0] - xe);
    ce = angle_coef_exact(*r.pr, rel - yaw);
    co = angle_coef_exact(*r.pr, 3.14159265358979323846 + rel - psi);
}

// ---------------------------------------------------------------------------------------------
// harm logits of one (ego sample t, prediction sample t) pair, the streaming form of phase 2
// ---------------------------------------------------------------------------------------------
// Coefficient of the geometric band of a direction (fp = P|P|, d2): first threshold whose test passes.  NT
// thresholds are compile-time; thresholds and coefficients are constant-bank operands (no loads, no loop).
// kSym: both sides share one coefficient table (the symmetric models), the side is not looked at.
// The reference's threshold sets are symmetric about pi/2 (45/135 and 15/45/75/105/135/165 degrees), i.e.
// fcos[NT-1-i] = -fcos[i]: the launcher checks this before it picks a specialisation.  Then the distance of fp to the
// nearest threshold is min_j | |fp| - d2 fcos[j] | over the NT/2 positive ones: one test per pair decides whether the
// sample is too close to call in fp32 (`amb`, see direction_band).
template <int NT, bool kSym>
__device__ __forceinline__ float chain_coef(const DevParams& pr, float fp, float d2, bool neg, float tol, bool& amb) {
    float coef = (!kSym && neg) ? pr.f_coef_neg[NT] : pr.f_coef_pos[NT];
    float h[NT / 2 > 0 ? NT / 2 : 1];
#pragma unroll
    for (int j = 0; j < NT / 2; ++j) h[j] = __fmul_rn(d2, pr.f_fcos[j]);
#pragma unroll
    for (int i = NT - 1; i >= 0; --i) {
        const float ci = (!kSym && neg) ? pr.f_coef_neg[i] : pr.f_coef_pos[i];
        const bool pass = i < NT / 2 ? fp > h[i] : fp > -h[NT - 1 - i];
        coef = pass ? ci : coef;
    }
    const float af = fabsf(fp);
#pragma unroll
    for (int j = 0; j < NT / 2; ++j) amb = amb || fabsf(__fsub_rn(af, h[j])) < tol;
    return coef;
}

constexpr float kWrapFree = 100.0f;  // marker in the ego "wrap" slot: un-wrapped and geometric ego angle bin alike

// R = float, NT >= 0: specialised.  `fast_ego`: every ego sample of the tile carries kWrapFree.
//
// Obstacle side (quirk Q4).  The reference bins a_o = pi + rel - psi_o with rel = atan2(dy, dx) in (-pi, pi] WITHOUT
// wrapping.  With theta = angle of the ego as seen from the obstacle's heading, a_o = pi + theta + 2 pi k for an
// integer k that depends only on the half plane of theta (sign of qo), on the sign of dy, on the sign of sin(psi_o)
// and on the whole turns in psi_o.  |a_o| <= pi (a band other than the last) needs k = 0 with theta <= 0 or k = -1
// with theta >= 0; the last two inputs are per prediction sample, so pack_obstacles_kernel tabulates the answer
// for the four (half plane, sign of dy) combinations in 4 bits of the obstacle record.
template <typename R, int NT, bool kSym = false> struct Logits {
    static __device__ __forceinline__ float dv(Vec4<float> c0, Vec2<float> c1, Vec4<float> g1) {
        return delta_v(c0.z, c0.w, c1.x, 0.0f, g1.x, g1.y, g1.z, 0.0f);
    }
    // impact-area coefficients of ego and obstacle for a protected obstacle; branch free
    static __device__ __forceinline__ void coefs(const DevParams& pr, const Tab<float>& tab, bool fast_ego, Vec4<float> c0,
                                                 Vec2<float> c1, Vec4<float> g0, Vec4<float> g1, float& ce, float& co,
                                                 bool& amb) {
        // every rounding below is spelled out (no contraction freedom): phase 2b repeats this test on the samples
        // phase 2a flagged and must find the same lanes
        const float dx = __fsub_rn(g0.x, c0.x), dy = __fsub_rn(g0.y, c0.y);
        const float d2 = __fmaf_rn(dx, dx, __fmul_rn(dy, dy));
        const float tol = band_tol(d2, dx, dy, tab.eps_pos);
        const float ec = c0.z, es = c0.w, ocp = g1.y, osp = g1.z;
        // ego: angle = rel - yaw_e
        const float p = __fmaf_rn(dx, ec, __fmul_rn(dy, es)), q = __fmaf_rn(dy, ec, -__fmul_rn(dx, es));
        ce = chain_coef<NT, kSym>(pr, __fmul_rn(p, fabsf(p)), d2, q < 0, tol, amb);
        if (!fast_ego) {
            int n0 = 0;
            if (dy >= 0 && es < 0 && q < 0) n0 = 1;
            if (dy < 0 && es >= 0 && q > 0) n0 = -1;
            const float ewrap = c1.y == kWrapFree ? 0.0f : c1.y;
            const float cl = q < 0 ? pr.f_coef_neg[NT] : pr.f_coef_pos[NT];
            ce = ((float)n0 != ewrap) ? cl : ce;
        }
        // obstacle: angle = pi + rel - psi_o
        const float po = __fmaf_rn(dx, ocp, __fmul_rn(dy, osp)), qo = __fmaf_rn(dy, ocp, -__fmul_rn(dx, osp));
        const bool pos = qo < 0 || (qo == 0 && po >= 0);
        const int code = __float_as_int(g1.w);  // validity table, stored as raw bits in the fp32 tile
        const bool valid = ((code >> ((pos ? 2 : 0) | (dy >= 0 ? 1 : 0))) & 1) != 0;
        co = chain_coef<NT, kSym>(pr, -__fmul_rn(po, fabsf(po)), d2, !pos, tol, amb);
        co = valid ? co : pr.f_coef_pos[NT];
    }
};

// generic form: runtime threshold count, any arithmetic type (fp64 check mode, unusual tables)
template <typename R, bool kSym> struct Logits<R, -1, kSym> {
    static __device__ __forceinline__ R dv(Vec4<R> c0, Vec2<R> c1, Vec4<R> g1) {
        const R ewrap = (sizeof(R) == 4 && c1.y == (R)kWrapFree) ? (R)0 : c1.y;
        return delta_v(c0.z, c0.w, c1.x, ewrap, g1.x, g1.y, g1.z, g1.w);
    }
    static __device__ __forceinline__ void coefs(const DevParams&, const Tab<R>& tab, bool, Vec4<R> c0, Vec2<R> c1, Vec4<R> g0,
                                                 Vec4<R> g1, R& ce, R& co, bool& amb) {
        const R ewrap = (sizeof(R) == 4 && c1.y == (R)kWrapFree) ? (R)0 : c1.y;
        impact_coefs(tab, c0.x, c0.y, c0.z, c0.w, ewrap, g0.x, g0.y, g1.y, g1.z, g1.w, ce, co, amb);
    }
};

// ---------------------------------------------------------------------------------------------
// collision probability of one (ego sample i = t+1, obstacle sample t) pair
// (collision_probability.py:205-252, 329-364): 3 means x 3 axis-aligned ego boxes.
// ---------------------------------------------------------------------------------------------
// diagnostics: evaluations of the fp32 path that took the fp64 fallback (|rho| > 0.55 or NaN)
__device__ unsigned long long g_fallback_evals = 0ull;

struct RectDouble {
    __device__ __forceinline__ double operator()(double xl, double xu, double yl, double yu, double r) const {
        return rect_prob(xl, xu, yl, yu, r);
    }
};
struct RectFallbackF {
    __device__ __forceinline__ float operator()(float xl, float xu, float yl, float yu, float r) const {
        return rect_prob_fallback_f(xl, xu, yl, yu, r);
    }
};
template <int N> struct RectNodesF {
    RhoNodes<N> n;
    __device__ __forceinline__ float operator()(float xl, float xu, float yl, float yu, float) const {
        return rect_prob_nodes_f<N>(n, xl, xu, yl, yu);
    }
};

// (ddx, ddy) = ego centre - obstacle mean.  Both positions are hundreds of metres from the origin, their difference
// is a few metres and the rectangle limits are differences divided by a standard deviation that can be 0.14 m: the
// caller forms (ddx, ddy) from split (high + low) fp32 coordinates so that the limits carry no cancellation error;
// everything added below (box offsets, obstacle half length, half extents) is metres-sized.
template <typename R, typename Rect>
__device__ __forceinline__ R nine_rectangles(const Tab<R>& pr, const Rect& rect, R ddx, R ddy, R c, R s, R ex, R ey, R isx,
                                             R isy, R rho) {
    const R hx = pr.hx, hy = pr.hy;
    const R ox = pr.rr * c, oy = pr.rr * s;  // offset of the outer boxes along the ego heading
    R prob = 0;
#pragma unroll 1
    for (int km = 0; km < 3; ++km) {
        const R sg = km == 0 ? (R)0 : (km == 1 ? (R)1 : (R)-1);
        const R ux = ddx - sg * ex, uy = ddy - sg * ey;  // ego centre - mean_km
#ifdef ETP_UNROLL_JB
#pragma unroll  // build variant: the three boxes of one mean as straight-line code (three copies of the rectangle body)
#else
#pragma unroll 1  // the rectangle body stays one small loop: unrolled, its nine copies thrash the instruction cache
#endif
        for (int jb = 0; jb < 3; ++jb) {
            const R sb = jb == 0 ? (R)0 : (jb == 1 ? (R)1 : (R)-1);
            const R bx = ux + sb * ox, by = uy + sb * oy;  // box centre - mean
            const R xl = (bx - hx) * isx, xu = (bx + hx) * isx;
            const R yl = (by - hy) * isy, yu = (by + hy) * isy;
            prob += rect(xl, xu, yl, yu, rho);
        }
    }
    return prob / 3;
}

// `c`, `s` = cos / sin of the ego yaw, `g` = (mx, my, ex, ey), `p` = (1/sx, 1/sy, rho, class), `nd` = the five node
// vectors of this (t, o) record.  All lanes of a warp share g, p and nd.
__device__ __forceinline__ double collision_prob(const Tab<double>& pr, double ddx, double ddy, double c, double s,
                                              Vec4<double> g, Vec4<double> p, const float4*) {
    return nine_rectangles<double>(pr, RectDouble{}, ddx, ddy, c, s, g.z, g.w, p.x, p.y, p.z);
}

__device__ __forceinline__ float collision_prob(const Tab<float>& pr, float ddx, float ddy, float c, float s, Vec4<float> g,
                                             Vec4<float> p, const float4* nd) {
    const int cls = (int)p.w;  // warp-uniform: Gauss-Legendre node class of |rho| (pack_obstacles_kernel)
    if (cls == RC_ZERO) {
        RectNodesF<0> r;
        return nine_rectangles<float>(pr, r, ddx, ddy, c, s, g.z, g.w, p.x, p.y, p.z);
    }
    if (cls == RC_FP64) {
        {   // diagnostics (etp_debug_counters): one atomic per warp -- cls is warp-uniform, the active lanes are counted by their leader
            const unsigned act = __activemask();
            if ((int)(threadIdx.x & 31) == __ffs(act) - 1) atomicAdd(&g_fallback_evals, (unsigned long long)__popc(act));
        }
        return nine_rectangles<float>(pr, RectFallbackF{}, ddx, ddy, c, s, g.z, g.w, p.x, p.y, p.z);
    }
    const float4 a = __ldg(nd), b = __ldg(nd + 1), cc = __ldg(nd + 2), d = __ldg(nd + 3), f = __ldg(nd + 4);
    // s[0..5] = a.xyzw b.xy | q[0..5] = b.zw cc.xyzw | c[0..5] = d.xyzw f.xy | f.zw = low parts of the mean
    if (cls == RC_N4) {
        RectNodesF<4> r;
        r.n.s[0] = a.x; r.n.s[1] = a.y; r.n.s[2] = a.z; r.n.s[3] = a.w;
        r.n.q[0] = b.z; r.n.q[1] = b.w; r.n.q[2] = cc.x; r.n.q[3] = cc.y;
        r.n.c[0] = d.x; r.n.c[1] = d.y; r.n.c[2] = d.z; r.n.c[3] = d.w;
        return nine_rectangles<float>(pr, r, ddx, ddy, c, s, g.z, g.w, p.x, p.y, p.z);
    }
    RectNodesF<6> r;
    r.n.s[0] = a.x; r.n.s[1] = a.y; r.n.s[2] = a.z; r.n.s[3] = a.w; r.n.s[4] = b.x; r.n.s[5] = b.y;
    r.n.q[0] = b.z; r.n.q[1] = b.w; r.n.q[2] = cc.x; r.n.q[3] = cc.y; r.n.q[4] = cc.z; r.n.q[5] = cc.w;
    r.n.c[0] = d.x; r.n.c[1] = d.y; r.n.c[2] = d.z; r.n.c[3] = d.w; r.n.c[4] = f.x; r.n.c[5] = f.y;
    return nine_rectangles<float>(pr, r, ddx, ddy, c, s, g.z, g.w, p.x, p.y, p.z);
}

// The one out-of-line call of phase 2b: rectangle probability of an un-gated sample and, for the rare lane whose
// band test fp32 could not decide (`amb`), the exact impact-area coefficients.  Keeping both behind a single call site
// keeps the register allocation of the calling loop tight.
template <typename R> struct ProbCoef {
    R prob, ce, co;
};
template <typename R>
__device__ __noinline__ ProbCoef<R> ungated_eval(const Tab<R>& tab, R ddx, R ddy, R c, R s, Vec4<R> g, Vec4<R> p, const float4* nd,
                                                 bool with_prob, bool amb, ExactRef xr, R ce, R co) {
    ProbCoef<R> r{(R)0, ce, co};
    if (with_prob) r.prob = collision_prob(tab, ddx, ddy, c, s, g, p, nd);
    if (amb) {
        double a, b;
        band_coefs_exact(xr, a, b);
        r.ce = (R)a;
        r.co = (R)b;
    }
    return r;
}

// The 5 m gate (collision_probability.py:183-203) repeated in fp64 from the exact inputs: the fp32 path calls
// this when its own distances are too close to the threshold to decide.  Ego sample `ti` is regenerated.
__device__ __noinline__ bool gate_exact(const DevStep* stp, bool given, int pl, int id, int ti, int o, int t) {
    const DevStep& st = *stp;
    double xe, ye;
    if (given) {
        xe = st.given[((size_t)ETP_F_X * st.Tmax + ti) * st.given_stride + pl];
        ye = st.given[((size_t)ETP_F_Y * st.Tmax + ti) * st.given_stride + pl];
    } else {
        const double* gp = st.prof + (size_t)pl * PF_COUNT * st.Tmax;
        const double* pm = st.prof_meta + (size_t)pl * PM_COUNT;
        const bool low = pm[PM_LOW] != 0.0;
        const Poly5 q = lateral_poly(st, low, st.d_list[id], pm[PM_TEND], pm[PM_DS]);
        const TrajPoint tp = traj_point(q, low, gp, st.Tmax, ti, st.dt);
        xe = tp.f[ETP_F_X];
        ye = tp.f[ETP_F_Y];
    }
    const double* rp = st.raw_pos + ((size_t)o * st.Tp + t) * 2;
    const double y1 = st.raw_yaw[(size_t)o * st.Tp + t + 1];
    const double len = 2 * st.obs_const[o * OC_COUNT + OC_HALFLEN];
    const double fx = cos(y1) * len / 2, fy = sin(y1) * len / 2;
    double ax = rp[0] - xe, ay = rp[1] - ye;
    double m = sqrt(ax * ax + ay * ay);
    ax = (rp[0] + fx) - xe; ay = (rp[1] + fy) - ye;
    m = fmin(m, sqrt(ax * ax + ay * ay));
    ax = (rp[0] - fx) - xe; ay = (rp[1] - fy) - ye;
    m = fmin(m, sqrt(ax * ax + ay * ay));
    return m > 5.0;
}

// ---------------------------------------------------------------------------------------------
// collision of the ego box with the predicted obstacle boxes (validity level 1, ETP_COLLISION_* in etp_b200.h;
// validity_checks.py:125-141,216-274, prediction_helpers.py:138-210)
// ---------------------------------------------------------------------------------------------
template <typename R> struct Box {
    R x, y, c, s, l, w;  // centre, local x axis (cos, sin), half extents
};

// Largest separating-axis margin of two boxes over the four face normals: > 0 <=> disjoint.
template <typename R> __device__ __forceinline__ R sat_margin(const Box<R>& a, const Box<R>& b) {
    const R dx = b.x - a.x, dy = b.y - a.y;
    const R c = abs_(a.c * b.c + a.s * b.s), s = abs_(a.s * b.c - a.c * b.s);
    const R m0 = abs_(dx * a.c + dy * a.s) - (a.l + b.l * c + b.w * s);
    const R m1 = abs_(dy * a.c - dx * a.s) - (a.w + b.l * s + b.w * c);
    const R m2 = abs_(dx * b.c + dy * b.s) - (b.l + a.l * c + a.w * s);
    const R m3 = abs_(dy * b.c - dx * b.s) - (b.w + a.l * s + a.w * c);
    return rmax_(rmax_(m0, m1), rmax_(m2, m3));
}

// The box in the frame of `a` that encloses `a` and the next box `b` of the same object (ETP_COLLISION_SWEPT).
template <typename R> __device__ __forceinline__ Box<R> hull_box(const Box<R>& a, const Box<R>& b) {
    const R dx = b.x - a.x, dy = b.y - a.y;
    const R ux = dx * a.c + dy * a.s, uy = dy * a.c - dx * a.s;
    const R c = abs_(a.c * b.c + a.s * b.s), s = abs_(a.s * b.c - a.c * b.s);
    const R ex = b.l * c + b.w * s, ey = b.l * s + b.w * c;
    const R xlo = ux - ex < -a.l ? ux - ex : -a.l, xhi = ux + ex > a.l ? ux + ex : a.l;
    const R ylo = uy - ey < -a.w ? uy - ey : -a.w, yhi = uy + ey > a.w ? uy + ey : a.w;
    const R mx = (R)0.5 * (xlo + xhi), my = (R)0.5 * (ylo + yhi);
    return Box<R>{a.x + mx * a.c - my * a.s, a.y + mx * a.s + my * a.c, a.c, a.s, (R)0.5 * (xhi - xlo), (R)0.5 * (yhi - ylo)};
}

__device__ __forceinline__ Box<double> ego_box_exact(const DevStep& st, const DevParams& pr, bool given, int pl, int id, int ti) {
    double xe, ye, yaw;
    if (given) {
        xe = st.given[((size_t)ETP_F_X * st.Tmax + ti) * st.given_stride + pl];
        ye = st.given[((size_t)ETP_F_Y * st.Tmax + ti) * st.given_stride + pl];
        yaw = st.given[((size_t)ETP_F_YAW * st.Tmax + ti) * st.given_stride + pl];
    } else {
        const double* gp = st.prof + (size_t)pl * PF_COUNT * st.Tmax;
        const double* pm = st.prof_meta + (size_t)pl * PM_COUNT;
        const bool low = pm[PM_LOW] != 0.0;
        const Poly5 q = lateral_poly(st, low, st.d_list[id], pm[PM_TEND], pm[PM_DS]);
        const TrajPoint tp = traj_point(q, low, gp, st.Tmax, ti, st.dt);
        xe = tp.f[ETP_F_X];
        ye = tp.f[ETP_F_Y];
        yaw = tp.f[ETP_F_YAW];
    }
    return Box<double>{xe, ye, cos(yaw), sin(yaw), pr.veh_l * 0.5, pr.veh_w * 0.5};
}

__device__ __forceinline__ Box<double> obstacle_box_exact(const DevStep& st, int o, int t) {
    const double* rp = st.raw_pos + ((size_t)o * st.Tp + t) * 2;
    const double psi = st.raw_yaw[(size_t)o * st.Tp + t];
    return Box<double>{rp[0], rp[1], cos(psi), sin(psi), st.obs_const[o * OC_COUNT + OC_HALFLEN], st.obs_const[o * OC_COUNT + OC_HALFWID]};
}

// One (ego box t + 1, prediction box t) test repeated in fp64 from the exact inputs: the fp32 path calls this when its
// own margin is too close to zero.  `hull_e` / `hull_o`: the object is swept over its next sample.
__device__ __noinline__ bool collide_exact(const DevStep* stp, const DevParams* prp, bool given, int pl, int id, int o, int t,
                                           bool hull_e, bool hull_o) {
    Box<double> e = ego_box_exact(*stp, *prp, given, pl, id, t + 1);
    if (hull_e) e = hull_box(e, ego_box_exact(*stp, *prp, given, pl, id, t + 2));
    Box<double> b = obstacle_box_exact(*stp, o, t);
    if (hull_o) b = hull_box(b, obstacle_box_exact(*stp, o, t + 1));
    return !(sat_margin(e, b) > 0.0);
}

// ---------------------------------------------------------------------------------------------
// shared-memory plan of one CTA
// ---------------------------------------------------------------------------------------------
enum { MX_EGO_RISK = 0, MX_OBST_RISK, MX_EGO_ARG, MX_OBST_ARG, MX_COUNT };  // running maxima over time, [MX][O][32]
enum { P1_V = 0, P1_D, P1_JL, P1_JQ, P1_TRAV, P1_XE, P1_YE, P1_COUNT };                   // phase-1 partial sums, [W][P1][32]
enum { P3_SUM_E = 0, P3_SUM_O, P3_SUM_ABS, P3_RESP, P3_MM, P3_MAX_O, P3_MM_LO, P3_MM_HI, P3_COUNT };  // phase-3 partials, [W][P3][32]

template <typename R> struct Smem {
    using U = typename Enc<R>::U;
    Vec4<R>* e0;    // [Ts][32] x, y, cos yaw, sin yaw
    Vec2<R>* e1;    // [Ts][32] v, yaw (fp64) or its wrap count / kWrapFree (fp32)
    __half2* el;    // [Ts][32] low parts of x, y scaled by 2^14: exact x = e0.x + el.x / 2^14 (fp32 mode; zero in fp64 mode)
    double* p3;     // [kWarps][P3_COUNT][32]  -- aliases e0: the ego samples are dead once phase 2 is over
    U* mx;          // [MX_COUNT][On][32]
    double* p1;     // [kWarps][P1_COUNT][32]
    int* ff;        // [kWarps][32] first failing curvature sample * 4 + kind
    unsigned* queue;  // [(Ts-1) * On] (t << 16 | o) pairs that have a lane inside the gate's reach
#ifdef ETP_TMA_STAGE
    Vec4<R>* stage;            // [kWarps][kOB][seg_len][2] G records of the item a warp works on (bulk-copied)
    unsigned long long* bar;   // [kWarps] mbarriers of the copies
    __host__ __device__ static int seg_len_of(int Ts, int O) {
        const int n_groups = (O + kOB - 1) / kOB;
        int n_seg = 1;
        if (n_groups > 0) {
            n_seg = (ETP_SEG_FACTOR * kWarps + n_groups - 1) / n_groups;
            const int max_seg = (Ts - 1 + 7) / 8;
            n_seg = n_seg > max_seg ? max_seg : n_seg;
            n_seg = n_seg < 1 ? 1 : n_seg;
        }
        return (Ts - 1 + n_seg - 1) / n_seg;
    }
    __host__ __device__ static size_t stage_bytes(int Ts, int O) {
        return (size_t)kWarps * kOB * seg_len_of(Ts, O) * 2 * sizeof(Vec4<R>) + kWarps * sizeof(unsigned long long);
    }
#endif
    __host__ __device__ static size_t ego_bytes(int Ts) {
        const size_t e = 32 * (size_t)Ts * (sizeof(Vec4<R>) + sizeof(Vec2<R>) + sizeof(__half2));
        const size_t p = (size_t)kWarps * P3_COUNT * 32 * sizeof(double);
        return e > p ? e : p;
    }
    __host__ __device__ static size_t bytes(int Ts, int O) {
        const size_t On = O > 0 ? O : 1;
        size_t b = ego_bytes(Ts) + 32 * (MX_COUNT * On * sizeof(U) + kWarps * P1_COUNT * sizeof(double) + kWarps * sizeof(int)) +
                   sizeof(unsigned) * (size_t)(Ts - 1) * On + sizeof(int) * On;
#ifdef ETP_TMA_STAGE
        b = ((b + 15) & ~size_t(15)) + stage_bytes(Ts, O);
#endif
        return b;
    }
    __device__ __forceinline__ Smem(unsigned char* base, int Ts, int O) {
        const size_t On = O > 0 ? O : 1;
        e0 = reinterpret_cast<Vec4<R>*>(base);
        e1 = reinterpret_cast<Vec2<R>*>(e0 + (size_t)Ts * 32);
        el = reinterpret_cast<__half2*>(e1 + (size_t)Ts * 32);
        p3 = reinterpret_cast<double*>(base);
        p1 = reinterpret_cast<double*>(base + ego_bytes(Ts));
        mx = reinterpret_cast<U*>(p1 + kWarps * P1_COUNT * 32);
        ff = reinterpret_cast<int*>(mx + MX_COUNT * On * 32);
        queue = reinterpret_cast<unsigned*>(ff + kWarps * 32);
#ifdef ETP_TMA_STAGE
        {
            const size_t used = (size_t)(reinterpret_cast<unsigned char*>(queue + (size_t)(Ts - 1) * On) + sizeof(int) * On - base);
            stage = reinterpret_cast<Vec4<R>*>(base + ((used + 15) & ~size_t(15)));
            bar = reinterpret_cast<unsigned long long*>(stage + (size_t)kWarps * kOB * seg_len_of(Ts, O) * 2);
        }
#endif
    }
};

// ---------------------------------------------------------------------------------------------
// phase 2a: one item = kOB obstacles of one class x one time segment
// ---------------------------------------------------------------------------------------------
#ifdef ETP_TMA_STAGE
// bulk asynchronous copy global -> shared through the TMA engine, completion counted on an mbarrier (PTX ISA: cp.async.bulk,
// mbarrier.*); the generic-proxy loads of the staged records follow the wait on the barrier
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(b))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "ETP_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra ETP_DONE;\n"
        "bra ETP_WAIT;\n"
        "ETP_DONE:\n"
        "}\n" ::"r"(smem_u32(b)),
        "r"(parity)
        : "memory");
}
#endif

template <typename R> struct Phase2a {
    const DevStep* st;
    const DevParams* pr;
    const Tab<R>* tab;
    const Smem<R>* sm;
    const Vec4<R>* G;
    const Vec4<R>* C;
    R* o_prob;
    R* o_harm;
    const int* perm;
    int* qn;
    size_t Cp, harm_plane, mx_plane;
    int O, Ts, Tl, cl, lane, seg_len, n_seg, pl, id;
    bool active, mahal, given;
#ifdef ETP_TMA_STAGE
    const Vec4<R>* Gom;       // obstacle-major G plane
    Vec4<R>* stage;           // this warp's staging buffer [kOB][seg_len][2]
    unsigned long long* bar;  // this warp's mbarrier
    unsigned* parity;         // its phase
#endif
};

// The four samples of a step are independent dependency chains in ONE basic block (loads are clamped instead of
// branched around, masks are selects), so the scheduler can interleave them: with 16 resident warps per SM the
// latency of a broadcast load or a MUFU op is otherwise exposed.
//   kProt    : the group holds protected obstacles (impact-angle terms); a mixed group at the class boundary runs
//              here too and selects per obstacle
//   kFastEgo : every ego sample of the tile is wrap-free (Logits::coefs)
//   kDense   : all 32 lanes carry live candidates of the same length, all kOB obstacles exist and their predictions
//              cover the segment: no masks, no clamps
//   kPure    : (build variant -DETP_PROT_PURE) every obstacle of the group is protected: no per-obstacle class select
template <typename R, int NT, bool kSym, bool kFastEgo, bool kProt, bool kDense, bool kPure = false>
__device__ __forceinline__ void phase2a_item(const Phase2a<R>& a, int g, int seg) {
    using U = typename Enc<R>::U;
    using L = Logits<R, NT, kSym>;
    const DevParams& pr = *a.pr;
    const Tab<R>& tab = *a.tab;
    const int O = a.O, lane = a.lane;
    const int t0 = seg * a.seg_len, t1 = min(t0 + a.seg_len, a.Ts - 1);
    int ok[kOB], n_pred[kOB];
    // element offset of the obstacle's row in a [.][O][Cp] plane (32 bits where the dispatcher has checked they suffice)
    typename std::conditional<kDense, unsigned, size_t>::type off[kOB];
    bool valid[kOB], prot[kOB];
    R ke[kOB], ko[kOB], far2[kOB], hm_e[kOB], hm_o[kOB];
#pragma unroll
    for (int k = 0; k < kOB; ++k) {
        valid[k] = kDense || g * kOB + k < O;
        ok[k] = a.perm[min(g * kOB + k, O - 1)];
        off[k] = (typename std::conditional<kDense, unsigned, size_t>::type)((size_t)ok[k] * a.Cp);
        const Vec4<R> cr = ldg4<R>(a.C + ok[k]);
        n_pred[k] = valid[k] ? __ldg(a.st->pred_len + ok[k]) : 0;
        prot[k] = kPure ? true : (cr.z != (R)0);
        ke[k] = cr.x; ko[k] = cr.y;
        const R reach = (R)5.05 + (R)a.st->obs_const[ok[k] * OC_COUNT + OC_HALFLEN];
        far2[k] = reach * reach;
        hm_e[k] = hm_o[k] = (R)-INFINITY;
    }
#ifdef ETP_TMA_STAGE
    if (kDense) {
        // the 4 x (t1 - t0) records of this item: one bulk copy per obstacle out of the obstacle-major plane
        const unsigned bytes = (unsigned)((t1 - t0) * 2 * sizeof(Vec4<R>));
        __syncwarp();  // the warp has finished reading the previous item's records
        if (lane == 0) {
            mbar_expect_tx(a.bar, kOB * bytes);
#pragma unroll
            for (int k = 0; k < kOB; ++k)
                bulk_g2s(a.stage + (size_t)k * a.seg_len * 2, a.Gom + 2 * ((size_t)ok[k] * a.st->Tp + t0), bytes, a.bar);
        }
        mbar_wait(a.bar, *a.parity);
        *a.parity ^= 1u;
    }
#endif
    R* pptr = a.o_prob ? a.o_prob + (size_t)t0 * O * a.Cp + a.cl : nullptr;
    R* hptr = a.o_harm ? a.o_harm + (size_t)t0 * O * a.Cp + a.cl : nullptr;
    const Vec4<R>* gt = a.G + 2 * (size_t)t0 * O;  // records of sample t
    Vec4<R> c0 = a.sm->e0[t0 * 32 + lane];
    Vec2<R> c1 = a.sm->e1[t0 * 32 + lane];
    for (int t = t0; t < t1; ++t) {
        const Vec4<R> n0 = a.sm->e0[(t + 1) * 32 + lane];
        const Vec2<R> n1 = a.sm->e1[(t + 1) * 32 + lane];
        const bool has = kDense || t < a.Tl - 1;  // this lane's candidate has harm sample t (and ego sample t+1)
        unsigned near_bits = 0, amb_bits = 0;
        R ae[kOB], ao[kOB];
#pragma unroll
        for (int k = 0; k < kOB; ++k) {
            // harm samples are t < min(T-1, Tp) (harm_estimation.py:234); past the prediction the last record is
            // re-read and masked out
#ifdef ETP_TMA_STAGE
            const Vec4<R>* gp = kDense ? a.stage + ((size_t)k * a.seg_len + (t - t0)) * 2
                                       : a.G + 2 * ((size_t)max(min(t, n_pred[k] - 1), 0) * O + ok[k]);
#else
            const Vec4<R>* gp = kDense ? gt + 2 * ok[k] : a.G + 2 * ((size_t)max(min(t, n_pred[k] - 1), 0) * O + ok[k]);
#endif
#ifdef ETP_TMA_STAGE
            const Vec4<R> g0 = kDense ? gp[0] : ldg4<R>(gp), g1 = kDense ? gp[1] : ldg4<R>(gp + 1);  // staged: shared-memory broadcast
#else
            const Vec4<R> g0 = ldg4<R>(gp), g1 = ldg4<R>(gp + 1);
#endif
            const R dv = L::dv(c0, c1, g1);
            const R dve = ke[k] * dv, dvo = ko[k] * dv;
            const bool hk = kDense || (has && t < n_pred[k]);
            bool use = hk;  // the sample enters the running maxima here
            if (kProt) {
                R ce, co;
                bool amb = false;
                L::coefs(pr, tab, kFastEgo, c0, c1, g0, g1, ce, co, amb);
                amb = amb && prot[k] && hk;
                amb_bits |= amb ? (1u << k) : 0u;
                use = hk && !amb;  // too close to call: redone exactly below / in phase 2b
                const R pe = tab.lr_speed * dve + ce, po = tab.lr_speed * dvo + co;
                const R ue = tab.lr1s_speed * dve, uo = tab.ped_speed * dvo;
                ae[k] = prot[k] ? pe : ue;  // a mixed group at the class boundary
                ao[k] = prot[k] ? po : uo;
            } else {
                ae[k] = tab.lr1s_speed * dve;
                ao[k] = tab.ped_speed * dvo;
            }
            hm_e[k] = use ? rmax_(hm_e[k], ae[k]) : hm_e[k];
            hm_o[k] = use ? rmax_(hm_o[k], ao[k]) : hm_o[k];
            // within 5 m + half length of the centre: the three-point gate has to look (phase 2b)
            const R dx = g0.x - n0.x, dy = g0.y - n0.y;
            const bool near = hk && (kDense || t + 1 < n_pred[k]) && (a.mahal || !(dx * dx + dy * dy > far2[k]));
            near_bits |= near ? (1u << k) : 0u;
        }
        const unsigned flags = __reduce_or_sync(kFull, near_bits | (amb_bits << kOB));
        // queued: within reach of the gate, or a band test some lane could not decide in fp32 (left out of the maxima
        // above, flagged, redone in fp64 in phase 2b)
        const unsigned ambq = kProt ? (flags >> kOB) : 0u;
        const unsigned queued = (flags | ambq) & ((1u << kOB) - 1u);
        if (queued && lane == 0) {
#pragma unroll
            for (int k = 0; k < kOB; ++k)
                if (queued & (1u << k))
                    a.sm->queue[atomicAdd(a.qn, 1)] = ((unsigned)t << 16) | (((ambq >> k) & 1u) << 15) | (unsigned)ok[k];
        }
        if (pptr && (kDense || a.active)) {
#pragma unroll
            for (int k = 0; k < kOB; ++k)
                if (valid[k] && !(queued & (1u << k))) pptr[off[k]] = (R)0;
        }
        if (hptr && a.active) {  // optional per-sample harm (get_harm, harm_estimation.py:202)
#pragma unroll
            for (int k = 0; k < kOB; ++k) {
                if (!valid[k]) continue;
                const bool hk = has && t < n_pred[k];
                hptr[off[k]] = hk ? sigmoid_harm<R>(prot[k] ? tab.lr_const : tab.lr1s_const, ae[k]) : (R)0;
                hptr[a.harm_plane + off[k]] = hk ? sigmoid_harm<R>(prot[k] ? tab.lr_const : tab.neg_ped_const, ao[k]) : (R)0;
            }
        }
        c0 = n0;
        c1 = n1;
        gt += 2 * (size_t)O;
        if (pptr) pptr += (size_t)O * a.Cp;
        if (hptr) hptr += (size_t)O * a.Cp;
    }
#pragma unroll
    for (int k = 0; k < kOB; ++k) {
        if (!valid[k]) continue;
        U* m = a.sm->mx + (size_t)ok[k] * 32 + lane;
        if (a.n_seg == 1) {
            m[MX_EGO_ARG * a.mx_plane] = Enc<R>::enc(hm_e[k]);
            m[MX_OBST_ARG * a.mx_plane] = Enc<R>::enc(hm_o[k]);
        } else {
            atomicMax(m + MX_EGO_ARG * a.mx_plane, Enc<R>::enc(hm_e[k]));
            atomicMax(m + MX_OBST_ARG * a.mx_plane, Enc<R>::enc(hm_o[k]));
        }
    }
}

// dispatch of one item to its specialisation
template <typename R, int NT, bool kSym>
__device__ __forceinline__ void phase2a_dispatch(const Phase2a<R>& a, int g, int seg, bool unprot, bool tile_fast,
                                                 bool tile_full, bool pure) {
    // every obstacle of the group exists and is predicted over the whole segment?
    bool full = tile_full && !a.o_harm && (g + 1) * kOB <= a.O;
    if (full) {
        const int t1 = min((seg + 1) * a.seg_len, a.Ts - 1);
#pragma unroll
        for (int k = 0; k < kOB; ++k) full = full && __ldg(a.st->pred_len + a.perm[g * kOB + k]) >= t1 + 1;
    }
    if (unprot) {
        if (full) phase2a_item<R, NT, kSym, false, false, true>(a, g, seg);
        else phase2a_item<R, NT, kSym, false, false, false>(a, g, seg);
    } else if (tile_fast) {
#ifdef ETP_PROT_PURE
        if (full && pure) phase2a_item<R, NT, kSym, true, true, true, true>(a, g, seg);
        else
#endif
        if (full) phase2a_item<R, NT, kSym, true, true, true>(a, g, seg);
        else phase2a_item<R, NT, kSym, true, true, false>(a, g, seg);
    } else {
        phase2a_item<R, NT, kSym, false, true, false>(a, g, seg);
    }
}

// ---------------------------------------------------------------------------------------------
// road boundary (etp_set_road_boundary): ego boxes against the triangles outside the road, fp64
// ---------------------------------------------------------------------------------------------
// Largest separating-axis margin of a box and a triangle over the two box axes and the three edge normals.
__device__ __forceinline__ double sat_margin_triangle(const Box<double>& e, const double* q) {
    double ux[3], uy[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double dx = q[2 * k] - e.x, dy = q[2 * k + 1] - e.y;
        ux[k] = dx * e.c + dy * e.s;  // vertex in the box frame
        uy[k] = dy * e.c - dx * e.s;
    }
    double m = fmax(fmax(fmin(ux[0], fmin(ux[1], ux[2])) - e.l, -fmax(ux[0], fmax(ux[1], ux[2])) - e.l),
                    fmax(fmin(uy[0], fmin(uy[1], uy[2])) - e.w, -fmax(uy[0], fmax(uy[1], uy[2])) - e.w));
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int j = (k + 1) % 3, i = (k + 2) % 3;
        double nx = -(uy[j] - uy[k]), ny = ux[j] - ux[k];  // normal of edge k -> j, in the box frame
        const double len = sqrt(nx * nx + ny * ny);
        if (!(len > 0.0)) continue;                        // degenerate edge
        nx /= len; ny /= len;
        const double a = nx * ux[k] + ny * uy[k];          // the edge's own offset; the third vertex lies on one side
        const double b = nx * ux[i] + ny * uy[i];
        const double r = e.l * fabs(nx) + e.w * fabs(ny);  // box radius along the normal (box centred at the origin)
        m = fmax(m, fmax(fmin(a, b) - r, -fmax(a, b) - r));
    }
    return m;
}

// fp32 form on coordinates relative to the grid origin, normals left un-normalised (no sqrt, no division): `hit` if no
// axis separates; `*sure` cleared when no axis separates by more than tol while one comes within tol of doing so
__device__ __forceinline__ bool triangle_hit_f(float cx, float cy, float c, float s, float l, float w, const float* q, float tol,
                                               bool* sure) {
    float ux[3], uy[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float dx = __ldg(q + 2 * k) - cx, dy = __ldg(q + 2 * k + 1) - cy;
        ux[k] = dx * c + dy * s;
        uy[k] = dy * c - dx * s;
    }
    const float m0 = fmaxf(fminf(ux[0], fminf(ux[1], ux[2])) - l, -fmaxf(ux[0], fmaxf(ux[1], ux[2])) - l);
    const float m1 = fmaxf(fminf(uy[0], fminf(uy[1], uy[2])) - w, -fmaxf(uy[0], fmaxf(uy[1], uy[2])) - w);
    bool sep = m0 > 0.0f || m1 > 0.0f, clear = m0 > tol || m1 > tol, maybe = m0 > -tol || m1 > -tol;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int j = (k + 1) % 3, i = (k + 2) % 3;
        const float nx = -(uy[j] - uy[k]), ny = ux[j] - ux[k];
        const float a = nx * ux[k] + ny * uy[k], b = nx * ux[i] + ny * uy[i];
        const float an = fabsf(nx) + fabsf(ny);  // >= |n|: a conservative scale for the tolerance
        const float m = fmaxf(fminf(a, b), -fmaxf(a, b)) - (l * fabsf(nx) + w * fabsf(ny));
        sep = sep || m > 0.0f;
        clear = clear || m > tol * an;
        maybe = maybe || m > -tol * an;
    }
    if (!clear && maybe) *sure = false;
    return !sep;
}

// Does the box touch a triangle of the boundary?  kExact: fp64 on the absolute coordinates (check mode, and the
// re-decision of the fp32 kernels); else fp32 first pass, `*sure` cleared when it cannot decide.
template <bool kExact>
__device__ __forceinline__ bool boundary_hit(const DevBoundary& b, const Box<double>& e, bool* sure) {
    const double rx = e.l * fabs(e.c) + e.w * fabs(e.s), ry = e.l * fabs(e.s) + e.w * fabs(e.c);
    const double gx = e.x - b.x0, gy = e.y - b.y0;
    const double fx0 = floor((gx - rx) * b.inv_cell), fx1 = floor((gx + rx) * b.inv_cell);
    const double fy0 = floor((gy - ry) * b.inv_cell), fy1 = floor((gy + ry) * b.inv_cell);
    if (fx1 < 0.0 || fy1 < 0.0 || fx0 >= (double)b.nx || fy0 >= (double)b.ny) return false;
    const int ix0 = fx0 < 0.0 ? 0 : (int)fx0, ix1 = fx1 >= (double)b.nx ? b.nx - 1 : (int)fx1;
    const int iy0 = fy0 < 0.0 ? 0 : (int)fy0, iy1 = fy1 >= (double)b.ny ? b.ny - 1 : (int)fy1;
    const float cx = (float)gx, cy = (float)gy, c = (float)e.c, s = (float)e.s, l = (float)e.l, w = (float)e.w;
    const float ax = (float)rx + b.tol_f, ay = (float)ry + b.tol_f;
    bool hit = false;
    for (int iy = iy0; iy <= iy1; ++iy)
        for (int ix = ix0; ix <= ix1; ++ix) {
            const int c0 = __ldg(b.cell_start + iy * b.nx + ix), c1 = __ldg(b.cell_start + iy * b.nx + ix + 1);
            for (int k = c0; k < c1; ++k) {
                const int ti = __ldg(b.cell_items + k);
                if (kExact) {
                    hit = hit || !(sat_margin_triangle(e, b.tri + 6 * (size_t)ti) > 0.0);
                } else {
                    const float* q = b.trif + 6 * (size_t)ti;
                    // bounding boxes first: most triangles of a cell are nowhere near the box
                    const float x0 = __ldg(q), x1 = __ldg(q + 2), x2 = __ldg(q + 4), y0 = __ldg(q + 1), y1 = __ldg(q + 3), y2 = __ldg(q + 5);
                    if (fminf(x0, fminf(x1, x2)) > cx + ax || fmaxf(x0, fmaxf(x1, x2)) < cx - ax ||
                        fminf(y0, fminf(y1, y2)) > cy + ay || fmaxf(y0, fmaxf(y1, y2)) < cy - ay)
                        continue;
                    hit = triangle_hit_f(cx, cy, c, s, l, w, q, b.tol_f, sure) || hit;
                }
            }
        }
    return hit;
}

// Road-boundary pass of one warp (lanes = candidates): first ego box i = warp, warp + stride, ... that touches the boundary.
// Box i of the collision object (create_collision_object, validity_checks.py:125-141) is sample i, or the hull of
// samples i and i + 1 in swept mode.  Poses come from the shared-memory ego samples; a box the fp32 pass cannot decide
// is rebuilt from the exact samples and tested in fp64.
template <typename R>
__device__ __noinline__ int boundary_pass(const DevStep* stp, const DevParams* prp, const Vec4<R>* e0, const __half2* el, int Ts, int Tl,
                                          int lane, int warp, int stride, bool given, int pl, int id) {
    const DevStep& st = *stp;
    const DevBoundary& b = st.boundary;
    const bool swept = b.check == ETP_COLLISION_SWEPT;
    const int n_box = swept ? Tl - 1 : Tl;
    const double hl = prp->veh_l * 0.5, hw = prp->veh_w * 0.5;
    int first = 0x7fffffff;
    for (int i = warp; i < n_box && i < first; i += stride) {
        const Vec4<R> a = e0[i * 32 + lane];
        const float2 lo = __half22float2(el[i * 32 + lane]);
        Box<double> e{st.origin_x + (double)a.x + (double)lo.x * (1.0 / kLoScale), st.origin_y + (double)a.y + (double)lo.y * (1.0 / kLoScale),
                      (double)a.z, (double)a.w, hl, hw};
        if (swept) {
            const Vec4<R> n = e0[(i + 1) * 32 + lane];
            const float2 ln = __half22float2(el[(i + 1) * 32 + lane]);
            e = hull_box(e, Box<double>{st.origin_x + (double)n.x + (double)ln.x * (1.0 / kLoScale),
                                        st.origin_y + (double)n.y + (double)ln.y * (1.0 / kLoScale), (double)n.z, (double)n.w, hl, hw});
        }
        bool sure = true, hit;
        if (sizeof(R) == 8) {
            hit = boundary_hit<true>(b, e, &sure);
        } else {
            hit = boundary_hit<false>(b, e, &sure);
            if (!sure) {
                e = ego_box_exact(st, *prp, given, pl, id, i);
                if (swept) e = hull_box(e, ego_box_exact(st, *prp, given, pl, id, i + 1));
                hit = boundary_hit<true>(b, e, &sure);
            }
        }
        if (hit) first = i;
    }
    return first;
}

// traj.v[i] of the first box that left the road (risk_costs.py:114-116), regenerated in fp64
__device__ __noinline__ double leave_velocity(const DevStep* stp, bool given, int pl, int id, int col, int i) {
    const DevStep& st = *stp;
    if (given) return st.given[((size_t)ETP_F_V * st.Tmax + i) * st.given_stride + col];
    const double* gp = st.prof + (size_t)pl * PF_COUNT * st.Tmax;
    const double* pm = st.prof_meta + (size_t)pl * PM_COUNT;
    const bool low = pm[PM_LOW] != 0.0;
    const Poly5 q = lateral_poly(st, low, st.d_list[id], pm[PM_TEND], pm[PM_DS]);
    return traj_point(q, low, gp, st.Tmax, i, st.dt).f[ETP_F_V];
}

// Goal bookkeeping of one trajectory sample (calc_trajectory_cost.py:553-571 reached_goal_state, :520-550
// goal_position_reached_and_left_again, :452 velocity_costs): bit 0 = goal state reached at this sample (position,
// velocity and orientation), bit 1 = sample outside the goal area, bit 2 = sample 0 inside the goal area.
__device__ __noinline__ int goal_sample_flags(const DevGoal* gp, double x, double y, double v, double yaw, int t) {
    const DevGoal& g = *gp;
    const bool in = !g.has_area || point_in_goal(g, g.poly_start, g.verts, g.circles, x, y);
    const bool vok = !g.has_vel || (g.v_lo < v && v < g.v_hi);                    // helper_functions.py:408-421
    const bool ins = g.o_lo < yaw && yaw < g.o_hi;
    const bool ook = !g.has_ori || (g.ori_invert ? !ins : ins);                    // check_orientation, :924-931
    return ((in && vok && ook) ? GF_REACHED : 0) | (in ? 0 : GF_OUTSIDE) | ((t == 0 && in) ? GF_START_INSIDE : 0);
}

// ---------------------------------------------------------------------------------------------
// thread-block cluster helpers (kBatch kernels of a step with few tiles: one tile per cluster, see score_kernel)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_nctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_id_x() { unsigned r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_count_x() { unsigned r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of `p` (own shared memory) in the shared memory of CTA `rank` of the cluster
__device__ __forceinline__ unsigned dsmem_addr(const void* p, unsigned rank) {
    unsigned a;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"((unsigned)__cvta_generic_to_shared(p)), "r"(rank));
    return a;
}
// plain loads and stores only: a 64-bit max is no native shared-memory atomic (the local one is a compare-and-swap loop), and
// with remote red.max pushes the fp64 check mode selected another candidate once in a dozen cycles of the closed-loop test
__device__ __forceinline__ unsigned dsmem_ld(unsigned a, unsigned) { unsigned v; asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ unsigned long long dsmem_ld(unsigned a, unsigned long long) { unsigned long long v; asm volatile("ld.shared::cluster.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void dsmem_st(unsigned a, unsigned v) { asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void dsmem_st(unsigned a, unsigned long long v) { asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }

// ---------------------------------------------------------------------------------------------
// score_kernel
// ---------------------------------------------------------------------------------------------
// A kBatch launch of few tiles (a closed-loop planning cycle: 1 .. 18 tiles on 148 SMs) runs as thread-block clusters, ONE
// TILE PER CLUSTER: every CTA of the cluster generates the tile's ego samples (phase 1, redundantly: it is the latency of a
// handful of time steps either way), the phase-2 items -- and with them the queue of collision-probability samples -- are
// dealt round-robin to the cluster's CTAs, each CTA's maxima are folded into CTA 0's through distributed shared memory, and
// CTA 0 alone runs phase 3.  csize == 1 (every other launch) is the plain kernel.
// kBatch: the grid walks the tiles of MANY planning steps (scenario sweep, etp_plan_batch): tile -> scenario through
// ba.tile_scn, the scenario's DevStep is copied into shared memory when a CTA moves on to another scenario, and everything
// derived from the step (shared-memory plan, obstacle order, item counts) is derived again.  Arg-min only.
template <typename R, int NT, bool kSym, bool kGiven, bool kBatch>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
score_kernel(const __grid_constant__ DevStep stp, const __grid_constant__ DevParams pr, DevOut out,
             unsigned int* __restrict__ counters, const BatchArgs ba) {
    using U = typename Enc<R>::U;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ DevStep s_st;  // kBatch: the current scenario's step
    const DevStep& st = *(kBatch ? &s_st : &stp);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // position in the thread-block cluster (kBatch launches of few tiles; 0 of 1 otherwise) and this warp's place among the
    // cluster's warps
    const int crank = kBatch ? (int)cluster_ctarank() : 0, csize = kBatch ? (int)cluster_nctarank() : 1;
    const int gwarp = warp * csize + crank, gwarps = kWarps * csize;
    // everything below is a function of the step: set once here, and again per scenario in a batch
    int Ts = 2, O = 0, On = 1, Cs = 0, n_tiles = kBatch ? ba.n_tiles : 0;
    size_t Cp = out.c_stride, harm_plane = 0, mx_plane = 32;
    bool small_planes = true;
    Smem<R> sm(smem_raw, 2, 0);
    Tile<R> tile(nullptr, 0, 0);
    int* s_perm = nullptr;  // [On] obstacles sorted by class
    int n_groups = 0, n_seg = 1, seg_len = 1, n_items = 0, chunk = 1, chunk0 = 1;
    R* o_traj = reinterpret_cast<R*>(out.traj);
    R* o_prob = reinterpret_cast<R*>(out.prob);
    R* o_red = reinterpret_cast<R*>(out.red);
    R* o_cost = reinterpret_cast<R*>(out.cost);
    R* o_harm = reinterpret_cast<R*>(out.harm);
    const bool mahal = pr.mahalanobis != 0;

    __shared__ Tab<R> tab;
    __shared__ int s_tile, s_next, s_item, s_qn, s_qpos;
    __shared__ int s_nprot;
    __shared__ int s_hit[32];  // lane's candidate collides with a predicted box (pr.collision_check)
    __shared__ int s_goal[32]; // GF_* flags of the lane's candidate (st.goal.on)
    __shared__ int s_leave[32]; // first ego box of the lane's candidate that touches the road boundary (st.boundary.check)
    auto derive = [&]() {
        Ts = st.Tmax;
        O = st.has_pred ? st.O : 0;
        On = O > 0 ? O : 1;
        sm = Smem<R>(smem_raw, Ts, O);
        Cs = st.n_local;  // candidates scored by this call: profiles p_begin + shard_index + k * shard_count
        if (!kBatch) n_tiles = (Cs + 31) / 32;
        tile = Tile<R>(st.obs, st.O, st.Tp);
        harm_plane = (size_t)(Ts - 1) * O * Cp;
        mx_plane = (size_t)On * 32;
        small_planes = (size_t)On * Cp < (size_t)1 << 32;  // row offsets of the [.][O][Cp] planes fit 32 bits
        s_perm = reinterpret_cast<int*>(sm.queue + (size_t)(Ts - 1) * On);
        // phase-2a items: obstacle groups of kOB x time segments, about three items per warp
        n_groups = (O + kOB - 1) / kOB;
        n_seg = 1;
        if (n_groups > 0) {
            n_seg = (ETP_SEG_FACTOR * gwarps + n_groups - 1) / n_groups;
            int min_len = csize > 1 ? (csize >= 8 ? 1 : 8 / csize) : 8;  // at least 8 samples per segment; a cluster cuts finer
            if (kBatch && csize > 1 && ba.dbg_min_len > 0) min_len = ba.dbg_min_len;
            const int max_seg = (Ts - 1 + min_len - 1) / min_len;
            n_seg = n_seg > max_seg ? max_seg : n_seg;
            n_seg = n_seg < 1 ? 1 : n_seg;
        }
        seg_len = (Ts - 1 + n_seg - 1) / n_seg;
        n_seg = seg_len > 0 ? (Ts - 1 + seg_len - 1) / seg_len : 1;  // no empty trailing segments
        n_seg = n_seg < 1 ? 1 : n_seg;
        n_items = n_groups * n_seg;
        // phase-1 time chunks: warp 0 takes about half a share, because it is the warp that still has the previous tile's
        // phase 3b (cost assembly, one candidate per lane) to finish while the others already generate
        chunk0 = ((Ts + kWarps - 1) / kWarps) / 2;
        chunk0 = chunk0 < 1 ? 1 : chunk0;
        chunk = (Ts - chunk0 + kWarps - 2) / (kWarps - 1);
        chunk = chunk < 1 ? 1 : chunk;
    };
    auto sort_obstacles = [&]() {  // every thread calls it; a barrier follows at the call sites
        if (tid == 0) fill_tab<R>(tab, pr, st.eps_pos);
        // protected obstacles first, each class in its original order: a phase-2a item then holds one class and needs no
        // per-sample class branch.  One warp, 32 obstacles per round: a ballot and a prefix count instead of a chain of
        // dependent loads in one thread (the order is the serial loop's).
        if (warp == kWarps - 1) {
            int n = 0;
            for (int pass = 0; pass < 2; ++pass) {
                for (int base = 0; base < O; base += 32) {
                    const int o = base + lane;
                    const bool prot_o = o < O && ldg4<R>(tile.C + o).z != (R)0;
                    const bool mine = o < O && (pass == 0 ? prot_o : !prot_o);
                    const unsigned m = __ballot_sync(kFull, mine);
                    if (mine) s_perm[n + __popc(m & ((1u << lane) - 1u))] = o;
                    n += __popc(m);
                }
                if (pass == 0 && lane == 0) s_nprot = n;
            }
        }
    };
    int cur_scn = -1;
#ifdef ETP_PHASE_CLOCK
    // debug variant: SM cycle count at the barriers of CTA 0's first tile, printed at the end (tools/build_variants.py)
    long long pclk[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int pclk_tiles = 0;
#define ETP_PCLK(k) do { if (tid == 0 && blockIdx.x == 0 && pclk[k] == 0) pclk[k] = clock64(); } while (0)
#else
#define ETP_PCLK(k) do { } while (0)
#endif
    ETP_PCLK(0);
    if (!kBatch) {
        derive();
        sort_obstacles();
    }
    ETP_PCLK(1);
#ifdef ETP_TMA_STAGE
    unsigned tma_parity = 0;  // phase of this warp's mbarrier (one barrier per warp, re-used item after item)
    if (!kBatch) {
        if (lane == 0) mbar_init(sm.bar + warp, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncwarp();
    }
#endif

    // the index of the next tile is fetched one tile ahead (by warp 1, at the start of phase 3a), so that the round trip of
    // the global atomic is off the serial path between two tiles
    // (a cluster takes the tiles cluster id, cluster id + clusters, ...: its CTAs must agree without talking)
    if (tid == 0) s_next = csize > 1 ? (int)cluster_id_x() : (int)atomicAdd(&counters[1], 1u);
    for (;;) {
        if (tid == 0) {
            s_tile = s_next;
            s_item = 0;
            s_qn = 0;
            s_qpos = 0;
        }
        if (tid < 32) {
            s_hit[tid] = 0;
            s_goal[tid] = 0;
            s_leave[tid] = 0x7fffffff;
        }
        for (int i = tid; i < MX_COUNT * On * 32; i += kThreads) sm.mx[i] = Enc<R>::enc((R)-INFINITY);
        __syncthreads();  // (a) tile index visible; previous tile fully consumed
        const int tl = s_tile;
        if (tl >= n_tiles) break;
#ifdef ETP_PHASE_CLOCK
        ++pclk_tiles;
#endif
        if (kBatch) {
            const int scn = __ldg(ba.tile_scn + tl);
            if (scn != cur_scn) {  // the same for every thread of the CTA; warp 0 is through with the previous step
                const int* src = reinterpret_cast<const int*>(ba.steps + scn);
                for (int i = tid; i < (int)(sizeof(DevStep) / sizeof(int)); i += kThreads) reinterpret_cast<int*>(&s_st)[i] = __ldg(src + i);
                __syncthreads();
                derive();
                sort_obstacles();
#ifdef ETP_TMA_STAGE
                if (lane == 0) mbar_init(sm.bar + warp, 1);  // the barriers move with the shared-memory plan
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
                tma_parity = 0;
#endif
                for (int i = tid; i < MX_COUNT * On * 32; i += kThreads) sm.mx[i] = Enc<R>::enc((R)-INFINITY);
                __syncthreads();
                cur_scn = scn;
            }
        }

        ETP_PCLK(2);
        // ---- this lane's candidate -----------------------------------------------------------------
        const int cl = (kBatch ? tl - __ldg(ba.tile_first + cur_scn) : tl) * 32 + lane;
        const bool active = cl < Cs;
        const int clc = active ? cl : Cs - 1;
        // generated candidates: row of the profile table and lateral target; given trajectories: none of that
        const int kl = kGiven ? 0 : clc / st.nd, id = kGiven ? 0 : clc - kl * st.nd;
        const int pl = kGiven ? clc : st.shard_index + kl * st.shard_count;
        const int c_dense = kGiven ? clc : (st.p_begin + pl) * st.nd + id;
        const double* gp = kGiven ? st.given + clc : st.prof + (size_t)pl * PF_COUNT * Ts;
        const double* pm = st.prof_meta + (size_t)pl * PM_COUNT;
        const double ds = kGiven ? 0.0 : pm[PM_DS];
        const bool low = kGiven ? false : pm[PM_LOW] != 0.0;
        const int T = kGiven ? st.given_T[clc] : (int)pm[PM_T];
        const double dT = kGiven ? 0.0 : st.d_list[id];
        const bool live = active && (kGiven || !(ds <= fabs(dT)));  // frenet_functions.py:333-334
        const int Tl = live ? T : 0;

        // ---- phase 1: generation + transform, warps over time chunks --------------------------------
        int wf = (sizeof(R) == 4 && NT >= 0) ? 1 : 0;
        {
            double sv = 0, sd = 0, jl = 0, jq = 0, trav = 0;
            int first_fail = 0x7fffffff, gflags = 0;
            Poly5 q{};
            if (!kGiven) q = lateral_poly(st, low, dT, pm[PM_TEND], ds);
            const int ta = warp == 0 ? 0 : min(chunk0 + (warp - 1) * chunk, Ts), tb = warp == 0 ? min(chunk0, Ts) : min(ta + chunk, Ts);
            double px = 0, py = 0;
            for (int t = ta; t < tb; ++t) {
                if (t < Tl) {
                    const TrajPoint tp = kGiven ? given_point(gp, Ts, st.given_stride, t) : traj_point(q, low, gp, Ts, t, st.dt);
                    // validity_checks.py:158; -2 velocity, -1 acceleration (ETP_RELAX_Q2_ACCELERATION only): the minimum wins
                    if (kGiven && (pr.relaxed & ETP_RELAX_Q2_ACCELERATION) && !(fabs(tp.f[ETP_F_S_DD]) <= pr.a_max))
                        first_fail = min(first_fail, -1);
                    if (kGiven && !(fabs(tp.f[ETP_F_S_D]) <= pr.v_max)) first_fail = -2;
                    if (o_traj) {
#pragma unroll
                        for (int f = 0; f < ETP_NUM_FIELDS; ++f) o_traj[((size_t)f * Ts + t) * Cp + cl] = (R)tp.f[f];
                    }
                    R w;
                    if (sizeof(R) == 4)
                        w = (tp.wrap == 0.0 && fabs(tp.f[ETP_F_YAW]) <= pr.wrapfree_yaw) ? (R)kWrapFree : (R)tp.wrap;
                    else
                        w = (R)tp.f[ETP_F_YAW];
                    const double xr = tp.f[ETP_F_X] - st.origin_x, yr = tp.f[ETP_F_Y] - st.origin_y;  // relative to the step origin
                    sm.e0[t * 32 + lane] = Vec4<R>{(R)xr, (R)yr, (R)tp.cos_yaw, (R)tp.sin_yaw};
                    sm.e1[t * 32 + lane] = Vec2<R>{(R)tp.f[ETP_F_V], w};
                    // rounding residue of the fp32 coordinates, at most half an ulp (3e-5 m at 500 m): kept to 11 bits
                    sm.el[t * 32 + lane] = __floats2half2_rn((float)((xr - (double)(R)xr) * kLoScale), (float)((yr - (double)(R)yr) * kLoScale));
                    if (sizeof(R) == 4 && w != (R)kWrapFree) wf = 0;
                    sv += tp.f[ETP_F_V];
                    sd += fabs(tp.f[ETP_F_D]);
                    jl += tp.f[ETP_F_S_DDD] * tp.f[ETP_F_S_DDD];
                    jq += tp.f[ETP_F_D_DDD] * tp.f[ETP_F_D_DDD];
                    const int cf = curvature_fail(pr, tp.f[ETP_F_V], tp.f[ETP_F_CURV]);
                    if (cf) first_fail = min(first_fail, t * 4 + cf);  // a velocity failure (-1) stays the minimum
                    if (st.goal.on) gflags |= goal_sample_flags(&st.goal, tp.f[ETP_F_X], tp.f[ETP_F_Y], tp.f[ETP_F_V], tp.f[ETP_F_YAW], t);
                    if (t > ta) {  // calc_trajectory_cost.py:739-757
                        const double dx = px - tp.f[ETP_F_X], dy = py - tp.f[ETP_F_Y];
                        trav += sqrt(dx * dx + dy * dy);
                    }
                    px = tp.f[ETP_F_X];
                    py = tp.f[ETP_F_Y];
                } else {
                    sm.e0[t * 32 + lane] = Vec4<R>{(R)0, (R)0, (R)1, (R)0};
                    sm.e1[t * 32 + lane] = Vec2<R>{(R)0, sizeof(R) == 4 ? (R)kWrapFree : (R)0};
                    sm.el[t * 32 + lane] = __floats2half2_rn(0.0f, 0.0f);
                    if (o_traj && active) {
#pragma unroll
                        for (int f = 0; f < ETP_NUM_FIELDS; ++f) o_traj[((size_t)f * Ts + t) * Cp + cl] = (R)0;
                    }
                }
            }
            double* p1 = sm.p1 + (size_t)warp * P1_COUNT * 32 + lane;
            p1[P1_V * 32] = sv; p1[P1_D * 32] = sd; p1[P1_JL * 32] = jl; p1[P1_JQ * 32] = jq; p1[P1_TRAV * 32] = trav;
            // last position of the candidate (calc_dist_to_goal_pos): held by the warp whose chunk ends the trajectory
            const bool last = Tl > ta && Tl <= tb;
            p1[P1_XE * 32] = last ? px : 0.0; p1[P1_YE * 32] = last ? py : 0.0;
            sm.ff[warp * 32 + lane] = first_fail;
            if (gflags) atomicOr(&s_goal[lane], gflags);
        }
        const bool tile_fast = __syncthreads_and(wf) != 0;  // (b) ego samples complete; all of them wrap-free?
        ETP_PCLK(3);
        // travelled distance (calc_trajectory_cost.py:739-757): the segment that enters this warp's time chunk, from the
        // neighbour's last sample in shared memory (split coordinates: exact to 1e-8 m) instead of a regenerated point
        if (warp > 0 && chunk0 + (warp - 1) * chunk < Tl) {
            const int t = chunk0 + (warp - 1) * chunk;
            const Vec4<R> a = sm.e0[(t - 1) * 32 + lane], b = sm.e0[t * 32 + lane];
            const float2 la = __half22float2(sm.el[(t - 1) * 32 + lane]), lb2 = __half22float2(sm.el[t * 32 + lane]);
            const double dx = ((double)a.x - (double)b.x) + (double)(la.x - lb2.x) * (1.0 / kLoScale);
            const double dy = ((double)a.y - (double)b.y) + (double)(la.y - lb2.y) * (1.0 / kLoScale);
            sm.p1[((size_t)warp * P1_COUNT + P1_TRAV) * 32 + lane] += sqrt(dx * dx + dy * dy);
        }
        const bool tile_full = __all_sync(kFull, live && T == Ts) && small_planes;

        // ---- phase 2a: harm logits of every (t, o) sample, straight-line; samples inside the gate's reach are queued
        // item = (kOB obstacles, time segment).  Sample t of the harm model pairs ego[t] with prediction[t]
        // (harm_estimation.py:313-332); sample t of the probability pairs ego[t+1] with mean[t], covariance[t],
        // orientation[t+1] (collision_probability.py:168-203).
        if (O > 0) {
            Phase2a<R> pa{&st, &pr, &tab, &sm, tile.G, tile.C, o_prob, o_harm, s_perm, &s_qn, Cp, harm_plane, mx_plane, O, Ts, Tl, cl, lane,
                          seg_len, n_seg, pl, id, active, mahal, kGiven
#ifdef ETP_TMA_STAGE
                          , tile.Gom, sm.stage + (size_t)warp * kOB * seg_len * 2, sm.bar + warp, &tma_parity
#endif
            };
            for (;;) {
                int item = 0;
                if (lane == 0) item = atomicAdd(&s_item, 1);
                item = __shfl_sync(kFull, item, 0);
                if (kBatch && csize > 1) item = item * csize + crank;  // the cluster's items go round-robin to its CTAs
                if (item >= n_items) break;
                const int g = item % n_groups, seg = item / n_groups;
                const bool unprot = g * kOB >= s_nprot;  // groups are cut from the class-sorted obstacle order
                phase2a_dispatch<R, NT, kSym>(pa, g, seg, unprot, tile_fast, tile_full, (g + 1) * kOB <= s_nprot);
            }
            __syncthreads();  // (b2) queue complete
            ETP_PCLK(4);

            // ---- phase 2b: queued samples: three-point gate, collision probability, risk ------------------
            const int qn = s_qn;
#ifdef ETP_P2B_PIPE
            int e_next = 0;
            if (lane == 0) e_next = atomicAdd(&s_qpos, 1);
#endif
            for (;;) {
#ifdef ETP_P2B_PIPE
                // the index of the entry after this one is on its way while this one is evaluated
                const int e = __shfl_sync(kFull, e_next, 0);
                if (e >= qn) break;
                if (lane == 0) e_next = atomicAdd(&s_qpos, 1);
#else
                int e = 0;
                if (lane == 0) e = atomicAdd(&s_qpos, 1);
                e = __shfl_sync(kFull, e, 0);
                if (e >= qn) break;
#endif
                const unsigned ent = sm.queue[e];
                const int t = (int)(ent >> 16), o = (int)(ent & 0x7fffu);
                const bool amb_entry = (ent >> 15) & 1u;
                const size_t vi = (size_t)t * O + o;
                const bool has = t < Tl - 1;
                const Vec4<R> n0 = sm.e0[(t + 1) * 32 + lane];
                const Vec4<R> g0 = ldg4<R>(tile.G + 2 * vi);
                // entries flagged for a band re-check may lie past the probability's range (t + 1 < Tp)
                const bool in_range = has && t + 1 < __ldg(st.pred_len + o);
                bool ungated = in_range;
                if (!mahal && in_range) {
                    const R x = n0.x, y = n0.y;
                    R dx = g0.x - x, dy = g0.y - y;
                    R d2 = dx * dx + dy * dy;
                    dx = (g0.x + g0.z) - x; dy = (g0.y + g0.w) - y;
                    d2 = fmin(d2, dx * dx + dy * dy);
                    dx = (g0.x - g0.z) - x; dy = (g0.y - g0.w) - y;
                    d2 = fmin(d2, dx * dx + dy * dy);
                    if (sizeof(R) == 8) {
                        ungated = !(sqrt_(d2) > (R)5.0);
                    } else {
                        ungated = !(d2 > (R)25.0);
                        if (abs_(d2 - (R)25) < (R)2e-3 + (R)1e-5 * (abs_(x) + abs_(y)))
                            ungated = !gate_exact(&st, kGiven, pl, id, t + 1, o, t);  // fp32 cannot decide: repeat in fp64
                    }
                }
                R prob = 0;
                if (ungated) {
                    const Vec4<R> c0 = sm.e0[t * 32 + lane];
                    const Vec2<R> c1 = sm.e1[t * 32 + lane];
                    const Vec4<R> g1 = ldg4<R>(tile.G + 2 * vi + 1);
                    const Vec4<R> cr = ldg4<R>(tile.C + o);
                    const bool prt = cr.z != (R)0;
                    // harm coefficients of sample t in fp32 (same code as phase 2a), `amb` if fp32 cannot decide a band
                    R ce = 0, co = 0;
                    bool amb = false;
                    if (prt) Logits<R, NT, kSym>::coefs(pr, tab, tile_fast, c0, c1, g0, g1, ce, co, amb);
                    // ego centre - mean from split coordinates (see nine_rectangles)
                    R ddx = n0.x - g0.x, ddy = n0.y - g0.y;
                    if (sizeof(R) == 4) {
                        const float2 lo = __half22float2(sm.el[(t + 1) * 32 + lane]);
                        const float4 ml = __ldg(tile.N + kNodeVecs * vi + 4);
                        ddx += (R)(lo.x * (1.0f / kLoScale) - ml.z);
                        ddy += (R)(lo.y * (1.0f / kLoScale) - ml.w);
                    }
                    const ProbCoef<R> pc = ungated_eval<R>(tab, ddx, ddy, n0.z, n0.w, g0, ldg4<R>(tile.P + 2 * vi),
                                                           tile.N + kNodeVecs * vi, !mahal, amb,
                                                           ExactRef{&st, &pr, kGiven, pl, id, t, o}, ce, co);
                    prob = pc.prob;
                    if (mahal) {
                        const Vec4<R> p1 = ldg4<R>(tile.P + 2 * vi + 1);
                        prob = inv_mahalanobis<R>(n0.x, n0.y, g0.x, g0.y, p1.x, p1.y, p1.z);
                    }
                    const R dv = Logits<R, NT, kSym>::dv(c0, c1, g1);
                    const R ae = prt ? tab.lr_speed * (cr.x * dv) + pc.ce : tab.lr1s_speed * (cr.x * dv);
                    const R ao = prt ? tab.lr_speed * (cr.y * dv) + pc.co : tab.ped_speed * (cr.y * dv);
#ifndef ETP_EXACT_SIGMOID_2B
                    // the risk carries the probability's 1e-5: two MUFU ops per sigmoid are accurate enough here (the harm
                    // maxima themselves take the full-precision form in phase 3a)
                    const R he = sigmoid_risk(prt ? tab.lr_const : tab.lr1s_const, ae);
                    const R ho = sigmoid_risk(prt ? tab.lr_const : tab.neg_ped_const, ao);
#else
                    const R he = sigmoid_harm<R>(prt ? tab.lr_const : tab.lr1s_const, ae);
                    const R ho = sigmoid_harm<R>(prt ? tab.lr_const : tab.neg_ped_const, ao);
#endif
                    U* m = sm.mx + (size_t)o * 32 + lane;
                    atomicMax(m + MX_EGO_RISK * mx_plane, Enc<R>::enc(he * prob));  // risk_costs.py:88-95
                    atomicMax(m + MX_OBST_RISK * mx_plane, Enc<R>::enc(ho * prob));
                }
                if (amb_entry) {
                    // phase 2a left this sample out of some lanes' running maxima: the same test (same code, same
                    // inputs, bit-identical outcome) finds those lanes again; their coefficients are redone in fp64
                    const Vec4<R> c0 = sm.e0[t * 32 + lane];
                    const Vec2<R> c1 = sm.e1[t * 32 + lane];
                    const Vec4<R> g1 = ldg4<R>(tile.G + 2 * vi + 1);
                    const Vec4<R> cr = ldg4<R>(tile.C + o);
                    R ce, co;
                    bool amb = false;
                    Logits<R, NT, kSym>::coefs(pr, tab, tile_fast, c0, c1, g0, g1, ce, co, amb);
                    if (amb && has && cr.z != (R)0 && t < __ldg(st.pred_len + o)) {
                        double ced, cod;
                        band_coefs_exact(ExactRef{&st, &pr, kGiven, pl, id, t, o}, ced, cod);
                        const R dv = Logits<R, NT, kSym>::dv(c0, c1, g1);
                        const R ae = tab.lr_speed * (cr.x * dv) + (R)ced, ao = tab.lr_speed * (cr.y * dv) + (R)cod;
                        U* m = sm.mx + (size_t)o * 32 + lane;
                        atomicMax(m + MX_EGO_ARG * mx_plane, Enc<R>::enc(ae));
                        atomicMax(m + MX_OBST_ARG * mx_plane, Enc<R>::enc(ao));
                        if (o_harm && active) {
                            o_harm[vi * Cp + cl] = sigmoid_harm<R>(tab.lr_const, ae);
                            o_harm[harm_plane + vi * Cp + cl] = sigmoid_harm<R>(tab.lr_const, ao);
                        }
                    }
                }
                if (o_prob && active) o_prob[vi * Cp + cl] = prob;
            }

            // ---- collision with the predicted boxes (optional): warps over (obstacle, time segment) items ----------
            // pair t: ego box of sample t + 1 (time step + t + 1) against prediction box t; swept mode replaces each by
            // the hull with its successor.  One distance test against a per-item reach rejects most pairs.
            if (pr.collision_check != ETP_COLLISION_HOST) {
                const bool swept = pr.collision_check == ETP_COLLISION_SWEPT;
                const R hl = (R)(pr.veh_l * 0.5), hw = (R)(pr.veh_w * 0.5);
                const R tol = (R)1e-4 + (R)16 * tab.eps_pos;
                R reach_e = sqrt_(hl * hl + hw * hw) + (R)1e-2;
                if (swept) {  // largest step of this candidate
                    R stp = 0;
                    for (int t = 0; t + 1 < Tl; ++t) {
                        const Vec4<R> a = sm.e0[t * 32 + lane], b = sm.e0[(t + 1) * 32 + lane];
                        stp = rmax_(stp, abs_(b.x - a.x) + abs_(b.y - a.y));
                    }
                    reach_e += stp;
                }
                const int c_seg = (2 * gwarps + O - 1) / O, c_len = (Ts - 1 + c_seg - 1) / c_seg;
                bool hit = false;
                for (int item = gwarp; item < O * c_seg; item += gwarps) {
                    const int o = item % O, t_lo = (item / O) * c_len, t_hi = min(t_lo + c_len, Ts - 1);
                    const int n_o = __ldg(st.pred_len + o);
                    const int plen = Tl < n_o ? Tl : n_o;  // prediction_helpers.py:168
                    const bool hull_o = swept && plen >= 2;
                    const R ol = (R)__ldg(st.obs_const + o * OC_COUNT + OC_HALFLEN), ow = (R)__ldg(st.obs_const + o * OC_COUNT + OC_HALFWID);
                    R reach = reach_e + sqrt_(ol * ol + ow * ow);
                    if (swept) reach += (R)__ldg(st.obs_const + o * OC_COUNT + OC_MAXSTEP);
                    const R reach2 = reach * reach;
                    const int t_end = min(t_hi, min(n_o, st.Tp));  // no prediction box past n_o
                    const Vec4<R>* Go = tile.G + 2 * (size_t)o;
                    // four pairs per round: the loads and distance tests are independent chains, the box test runs only
                    // when one of them is within reach (a clamped tail pair is simply tested twice)
                    for (int t4 = t_lo; t4 < t_end; t4 += 4) {
                        unsigned near = 0;
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int t = min(t4 + k, t_end - 1);
                            const Vec4<R> n0 = sm.e0[(t + 1) * 32 + lane];
                            const Vec4<R> g0 = ldg4<R>(Go + 2 * (size_t)(t * O));
                            const R dx = g0.x - n0.x, dy = g0.y - n0.y;
                            near |= (dx * dx + dy * dy <= reach2 ? 1u : 0u) << k;
                        }
                        if (near == 0) continue;
#pragma unroll 1
                        for (int k = 0; k < 4; ++k) {
                            const int t = min(t4 + k, t_end - 1);
                            const bool valid = swept ? (t + 2 < Tl && (hull_o ? t + 1 < plen : t == 0)) : (t + 1 < Tl && t < plen);
                            if (!((near >> k) & 1u) || !valid) continue;
                            const Vec4<R> n0 = sm.e0[(t + 1) * 32 + lane];
                            const Vec4<R> g0 = ldg4<R>(Go + 2 * (size_t)(t * O)), g1 = ldg4<R>(Go + 2 * (size_t)(t * O) + 1);
                            Box<R> e{n0.x, n0.y, n0.z, n0.w, hl, hw};
                            Box<R> b{g0.x, g0.y, g1.y, g1.z, ol, ow};
                            if (swept) {
                                const Vec4<R> n1 = sm.e0[(t + 2) * 32 + lane];
                                e = hull_box(e, Box<R>{n1.x, n1.y, n1.z, n1.w, hl, hw});
                                if (hull_o) {
                                    const Vec4<R> h0 = ldg4<R>(Go + 2 * (size_t)((t + 1) * O)), h1 = ldg4<R>(Go + 2 * (size_t)((t + 1) * O) + 1);
                                    b = hull_box(b, Box<R>{h0.x, h0.y, h1.y, h1.z, ol, ow});
                                }
                            }
                            const R m = sat_margin(e, b);
                            bool col = !(m > (R)0);
                            if (sizeof(R) == 4 && abs_(m) < tol) col = collide_exact(&st, &pr, kGiven, pl, id, o, t, swept, hull_o);
                            hit |= col;
                        }
                    }
                    if (__all_sync(kFull, hit || !live)) break;  // nothing left to find for this warp's lanes
                }
                if (hit) s_hit[lane] = 1;
            }
        }
        // ---- road boundary (optional, etp_set_road_boundary): first ego box that touches it ---------------------
        if (st.boundary.check != ETP_COLLISION_HOST) {
            const int first = boundary_pass<R>(&st, &pr, sm.e0, sm.el, Ts, Tl, lane, gwarp, gwarps, kGiven, pl, id);
            if (first != 0x7fffffff) atomicMin(&s_leave[lane], first);
        }
        __syncthreads();  // (c) maxima complete, ego samples dead
        ETP_PCLK(5);
        if (tid == 32) s_next = csize > 1 ? s_tile + (int)cluster_count_x() : (int)atomicAdd(&counters[1], 1u);  // read by thread 0 after barrier (d)
        if (kBatch && csize > 1) {
            // every CTA of the cluster holds the maxima of its share of the samples: folded into CTA 0's planes, which then
            // assembles the costs alone
            cluster_sync_all();  // every CTA's maxima are complete
            // word i of the planes is folded by CTA (i / kThreads) % csize: it reads the word from every CTA and leaves the
            // maximum in CTA 0 -- plain distributed-shared-memory loads and one store, every word touched by one thread only
            for (int i = crank * kThreads + tid; i < MX_COUNT * On * 32; i += kThreads * csize) {
                U v = sm.mx[i];
                for (int r = 0; r < csize; ++r) {
                    if (r == crank) continue;
                    const U w = dsmem_ld(dsmem_addr(sm.mx + i, (unsigned)r), U{});
                    v = w > v ? w : v;
                }
                if (crank == 0) sm.mx[i] = v;
                else dsmem_st(dsmem_addr(sm.mx + i, 0u), v);
            }
            if (crank == 0 && tid < 32) {
                int hit = s_hit[tid], leave = s_leave[tid];
                for (int r = 1; r < csize; ++r) {
                    hit |= (int)dsmem_ld(dsmem_addr(&s_hit[tid], (unsigned)r), 0u);
                    leave = min(leave, (int)dsmem_ld(dsmem_addr(&s_leave[tid], (unsigned)r), 0u));
                }
                s_hit[tid] = hit;
                s_leave[tid] = leave;
            }
            cluster_sync_all();  // folded values in CTA 0; nobody touches another CTA's shared memory past this point
            if (crank != 0) continue;
        }

        // ---- phase 3a: per-obstacle reductions, warps over obstacles ----------------------------------
        {
            double sum_e = 0, sum_o = 0, sum_abs = 0, resp = 0, mm = -INFINITY, max_o = -INFINITY;
            double mm_lo = -INFINITY, mm_hi = -INFINITY;  // the gated harm maximum if every risk near eps fell outside / inside the gate
            for (int o = warp; o < O; o += kWarps) {
                R re = 0, ro = 0, he = 0, ho = 0;
                if (live) {
                    const Vec4<R> cr = ldg4<R>(tile.C + o);
                    const bool prt = cr.z != (R)0;
                    const U* m = sm.mx + (size_t)o * 32 + lane;
                    // every gated harm sample contributes harm * 0.0 to the maxima (risk_costs.py:88-101)
                    re = rmax_(Enc<R>::dec(m[MX_EGO_RISK * mx_plane]), (R)0);
                    ro = rmax_(Enc<R>::dec(m[MX_OBST_RISK * mx_plane]), (R)0);
                    he = sigmoid_harm<R>(prt ? tab.lr_const : tab.lr1s_const, Enc<R>::dec(m[MX_EGO_ARG * mx_plane]));
                    ho = sigmoid_harm<R>(prt ? tab.lr_const : tab.neg_ped_const, Enc<R>::dec(m[MX_OBST_ARG * mx_plane]));
                    sum_e += (double)re;
                    sum_o += (double)ro;
                    sum_abs += fabs((double)re - (double)ro);
                    resp -= st.obs_const[o * OC_COUNT + OC_RESP] * (double)ro;       // risk_costs.py:250-253
                    // risk_costs.py:205-206 keeps the harm where the risk is BELOW eps (Q3); ETP_RELAX_Q3_MAXIMIN turns the gate
                    const bool gate = (pr.relaxed & ETP_RELAX_Q3_MAXIMIN) != 0;
                    mm = fmax(mm, (double)he * ((((double)re < 10e-10) != gate) ? 1.0 : 0.0));
                    mm = fmax(mm, (double)ho * ((((double)ro < 10e-10) != gate) ? 1.0 : 0.0));
                    max_o = fmax(max_o, (double)ro);
                    // a risk within kTolGate of the gate's eps: fp32 cannot say on which side it lies (st.refine); it matters
                    // only if the two readings give different maxima
                    const double e_in = 10e-10 * (1.0 - kTolGate), e_out = 10e-10 * (1.0 + kTolGate);
                    mm_lo = fmax(mm_lo, (double)he * ((((double)re < (gate ? e_out : e_in)) != gate) ? 1.0 : 0.0));
                    mm_lo = fmax(mm_lo, (double)ho * ((((double)ro < (gate ? e_out : e_in)) != gate) ? 1.0 : 0.0));
                    mm_hi = fmax(mm_hi, (double)he * ((((double)re < (gate ? e_in : e_out)) != gate) ? 1.0 : 0.0));
                    mm_hi = fmax(mm_hi, (double)ho * ((((double)ro < (gate ? e_in : e_out)) != gate) ? 1.0 : 0.0));
                }
                if (o_red && active) {
                    o_red[((size_t)ETP_R_EGO_RISK * O + o) * Cp + cl] = re;
                    o_red[((size_t)ETP_R_OBST_RISK * O + o) * Cp + cl] = ro;
                    o_red[((size_t)ETP_R_EGO_HARM * O + o) * Cp + cl] = he;
                    o_red[((size_t)ETP_R_OBST_HARM * O + o) * Cp + cl] = ho;
                }
            }
            double* p3 = sm.p3 + (size_t)warp * P3_COUNT * 32 + lane;
            p3[P3_SUM_E * 32] = sum_e; p3[P3_SUM_O * 32] = sum_o; p3[P3_SUM_ABS * 32] = sum_abs;
            p3[P3_RESP * 32] = resp; p3[P3_MM * 32] = mm; p3[P3_MAX_O * 32] = max_o; p3[P3_MM_LO * 32] = mm_lo; p3[P3_MM_HI * 32] = mm_hi;
        }
        __syncthreads();  // (d) partials complete
        ETP_PCLK(6);

        // ---- phase 3b: principles, cost, level (warp 0, one candidate per lane) -------------------------
        if (warp == 0) {
            double c[ETP_NUM_COST];
#pragma unroll
            for (int k = 0; k < ETP_NUM_COST; ++k) c[k] = 0.0;
            int level = ETP_LEVEL_SKIPPED, reason = ETP_REASON_NONE;
            double bd_out = 0.0;
            float bound = 0.0f;
            bool amb = false;
            if (live) {
                CandFacts f;
                RiskSums r{0, 0, 0, 0, -INFINITY, -INFINITY};
                f.sv = f.sd = f.jl = f.jq = f.trav = f.x_end = f.y_end = 0.0;
                double mm_lo = -INFINITY, mm_hi = -INFINITY;
                int first_fail = 0x7fffffff;
                for (int w = 0; w < kWarps; ++w) {
                    const double* p1 = sm.p1 + (size_t)w * P1_COUNT * 32 + lane;
                    const double* p3 = sm.p3 + (size_t)w * P3_COUNT * 32 + lane;
                    f.sv += p1[P1_V * 32]; f.sd += p1[P1_D * 32]; f.jl += p1[P1_JL * 32]; f.jq += p1[P1_JQ * 32];
                    f.trav += p1[P1_TRAV * 32];
                    f.x_end += p1[P1_XE * 32]; f.y_end += p1[P1_YE * 32];
                    r.sum_e += p3[P3_SUM_E * 32]; r.sum_o += p3[P3_SUM_O * 32]; r.sum_abs += p3[P3_SUM_ABS * 32];
                    r.resp += p3[P3_RESP * 32];
                    r.mm = fmax(r.mm, p3[P3_MM * 32]);
                    r.max_o = fmax(r.max_o, p3[P3_MAX_O * 32]);
                    mm_lo = fmax(mm_lo, p3[P3_MM_LO * 32]);
                    mm_hi = fmax(mm_hi, p3[P3_MM_HI * 32]);
                    first_fail = min(first_fail, sm.ff[w * 32 + lane]);
                }
                f.T = T;
                const bool vel_ok = kGiven ? first_fail != -2 : pm[PM_VEL_OK] != 0.0;
                const bool acc_ok = kGiven ? first_fail != -1 : !(pm[PM_VEL_OK] < 0.0);  // fails only under ETP_RELAX_Q2_ACCELERATION
                double bd = st.bd_harm ? st.bd_harm[c_dense] : 0.0;
                int ext = st.ext_level ? (int)st.ext_level[c_dense] : 10;
                if (st.boundary.check != ETP_COLLISION_HOST) {
                    // risk_costs.py:104-123: harm of leaving the road at the speed of the first colliding box; without
                    // predictions calc_risk does not run and boundary_valid alone decides (validity_checks.py:108-112)
                    const int leave = s_leave[lane];
                    bd = 0.0;
                    if (leave != 0x7fffffff) {
                        const double vl = leave_velocity(&st, kGiven, pl, id, clc, leave);
                        bd = 1.0 / (1.0 + exp(-st.boundary.lr1s_const - st.boundary.lr1s_speed * vl));
                        if (ext != 1) ext = 2;
                    } else if (ext == 2) {
                        ext = 10;
                    }
                }
                bd_out = bd;
                f.bd = bd;
                f.gf = s_goal[lane];
                f.factor = st.factor_list ? st.factor_list[c_dense] : st.factor;
                // validity tests that precede the risk (validity_checks.py:31-115)
                const int cf = (first_fail == 0x7fffffff || first_fail < 0) ? 0 : (first_fail & 3);
                f.pre_level = -1; f.pre_reason = ETP_REASON_NONE;
                if (!vel_ok) { f.pre_level = 0; f.pre_reason = ETP_REASON_VELOCITY; }
                else if (!acc_ok) { f.pre_level = 0; f.pre_reason = ETP_REASON_ACCELERATION; }
                else if (cf) { f.pre_level = 0; f.pre_reason = cf == 1 ? ETP_REASON_CURV_LVC : ETP_REASON_CURV_HVC; }
                else if (ext == 1 || s_hit[lane]) { f.pre_level = 1; f.pre_reason = ETP_REASON_COLLISION; }
                else if (st.has_pred ? (bd != 0.0) : (ext == 2)) { f.pre_level = 2; f.pre_reason = ETP_REASON_BOUNDARIES; }
                assemble_costs(st, pr, O, f, r, c, level, reason);
                if (st.refine) {
                    const bool risk_in_cost = pr.w[ETP_W_RISK_COST] > 0.0 && st.has_pred && pr.planner_mode == ETP_MODE_RISK;
                    bound = cost_bound(pr, O, r, bd, c, risk_in_cost);
                    // decisions fp32 cannot make: the maximin gate of an obstacle, the max-risk level
                    amb = O > 0 && fmax(mm_lo, bd) != fmax(mm_hi, bd);
                    if (f.pre_level < 0 && pr.planner_mode == ETP_MODE_RISK && st.has_pred && O > 0) {
                        const double tol = 2.0 * kTolRisk * pr.max_risk + O * kTolRiskAbs;
                        amb = amb || fabs(r.sum_e - pr.max_risk) <= tol || fabs(r.max_o - pr.max_risk) <= tol;
                    }
                }
            }
            if (active) {
                if (o_cost) {
#pragma unroll
                    for (int k = 0; k < ETP_NUM_COST; ++k) o_cost[(size_t)k * Cp + cl] = (R)c[k];
                }
                if (out.bd_harm) out.bd_harm[cl] = bd_out;
                if (out.level) out.level[cl] = (uint8_t)level;
                if (out.reason) out.reason[cl] = (uint8_t)reason;
                // selection workspace (etp_select.cuh)
                st.ws_cost[cl] = c[ETP_C_TOTAL];
                st.ws_meta[cl] = live ? (uint8_t)(level | (amb ? SEL_AMB : 0)) : (uint8_t)SEL_SKIPPED;
                if (st.refine) {
                    st.ws_bound[cl] = bound;
                    st.ws_bd[cl] = bd_out;
                    st.ws_gf[cl] = (uint8_t)s_goal[lane];
                }
            }
        }
        ETP_PCLK(7);
        // no barrier here: warp 0 reads only p1 / p3 / ff; p1 / ff are next written after barrier (a) of the next
        // tile, p3 (aliasing the ego samples) after barrier (a) as well
    }
    if (kBatch && ba.sel) {
        // planning cycle: the CTA that finishes last has every candidate's workspace entry in front of it and runs the
        // selection right here (one block of kSelThreads == kThreads threads) instead of in a launch of its own
        static_assert(kSelThreads == kThreads, "the fused kernel's CTA runs select_body as it is");
        __shared__ int s_closing;
        __syncthreads();  // this CTA's workspace entries are written (warp 0)
        if (tid == 0) {
            __threadfence();
            s_closing = atomicAdd(ba.done, 1u) == gridDim.x - 1;
        }
        __syncthreads();
        if (s_closing) {
            __threadfence();
            const int* src = reinterpret_cast<const int*>(ba.steps);
            for (int i = tid; i < (int)(sizeof(DevStep) / sizeof(int)); i += kThreads) reinterpret_cast<int*>(&s_st)[i] = __ldg(src + i);
            __shared__ SelectBufs s_sb;
            if (tid == 0) {
                s_sb = *reinterpret_cast<const SelectBufs*>(ba.sel);
                *ba.done = 0;
            }
            __syncthreads();
            // more candidates than one selection block holds: this CTA walks the blocks one after the other (the records
            // the blocks exchange go through global memory either way)
            const int nblk = (s_st.n_local + kSelChunk - 1) / kSelChunk;
            for (int b = 0; b < (nblk > 0 ? nblk : 1); ++b) {
                select_body(s_st, s_sb, ba.best, s_st.nskip, &counters[1], 0, 0, b, nblk > 0 ? nblk : 1);
                __syncthreads();
            }
            if (nblk > 1 && s_st.refine)
                for (int b = 0; b < nblk; ++b) {
                    select_collect_body(s_st, s_sb, ba.best, s_st.nskip, &counters[1], 0, 0, b, nblk);
                    __syncthreads();
                }
        }
    }
#ifdef ETP_PHASE_CLOCK
    if (tid == 0 && blockIdx.x == 0)
        printf("score_kernel CTA 0: %d tiles; cycles setup %lld, to tile start %lld, phase 1 %lld, phase 2a %lld, phase 2b %lld, phase 3a %lld, "
               "phase 3b %lld, whole %lld\n", pclk_tiles, pclk[1] - pclk[0], pclk[2] - pclk[1], pclk[3] - pclk[2], pclk[4] - pclk[3],
               pclk[5] - pclk[4], pclk[6] - pclk[5], pclk[7] - pclk[6], clock64() - pclk[0]);
#endif
}

// ---------------------------------------------------------------------------------------------
// host-side launcher
// ---------------------------------------------------------------------------------------------
// Resident CTAs per SM of one specialisation for a dynamic shared-memory size, asked of the runtime once and remembered
// (the attribute call and the occupancy query cost several microseconds each, which a 50 us planning step notices).
template <typename R, int NT, bool kSym, bool kGiven, bool kBatch>
static cudaError_t resident_ctas(size_t smem, int* per_sm_out) {
    constexpr int kSlots = 8, kDevices = 16;
    struct Seen {
        size_t smem[kSlots];
        int per_sm[kSlots];
        int n;
        size_t attr_smem;
    };
    static Seen seen[kDevices] = {};
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    Seen local{};
    Seen& sn = (dev >= 0 && dev < kDevices) ? seen[dev] : local;
    for (int i = 0; i < sn.n; ++i)
        if (sn.smem[i] == smem) { *per_sm_out = sn.per_sm[i]; return cudaSuccess; }
    if (smem > sn.attr_smem) {
        e = cudaFuncSetAttribute(score_kernel<R, NT, kSym, kGiven, kBatch>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        sn.attr_smem = smem;
    }
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, score_kernel<R, NT, kSym, kGiven, kBatch>, kThreads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    const int slot = sn.n < kSlots ? sn.n++ : 0;
    sn.smem[slot] = smem;
    sn.per_sm[slot] = per_sm;
    *per_sm_out = per_sm;
    return cudaSuccess;
}

template <typename R, int NT, bool kSym, bool kGiven, bool kBatch = false>
static cudaError_t launch_score_t(const DevStep& st, const DevParams& pr, const DevOut& out, unsigned* counters, int num_sms,
                                  int* grid_out, const BatchArgs* ba, size_t batch_smem, cudaStream_t stream) {
    const int O = st.has_pred ? st.O : 0;
    const size_t smem = kBatch ? batch_smem : Smem<R>::bytes(st.Tmax, O);
    int per_sm = 0;
    cudaError_t e = resident_ctas<R, NT, kSym, kGiven, kBatch>(smem, &per_sm);
    if (e != cudaSuccess) return e;
    // persistent grid: every SM fully occupied, never more CTAs than there are tiles
    const long long tiles = kBatch ? ba->n_tiles : ((long long)st.n_local + 31) / 32;
    long long grid = (long long)num_sms * per_sm;
    if (grid > tiles) grid = tiles;
    if (grid > kMaxGrid) grid = kMaxGrid;
    if (grid < 1) grid = 1;
    *grid_out = (int)grid;
    if (kBatch) {
        // few tiles (a planning cycle, not a sweep): one tile per thread-block cluster, as many CTAs per tile as leave every
        // CTA an SM of its own
        int csize = 1;
        static const bool no_cluster = getenv("ETP_NO_CLUSTER") != nullptr;
        static const int max_cluster = getenv("ETP_CLUSTER_MAX") ? atoi(getenv("ETP_CLUSTER_MAX")) : 8;
        static const int dbg_min_len = getenv("ETP_CLUSTER_MINLEN") ? atoi(getenv("ETP_CLUSTER_MINLEN")) : 0;
        static bool cluster_refused = false;  // a cluster launch failed once on this device: plain launches from then on
        for (int k = 8; k > 1 && !no_cluster && !cluster_refused; k >>= 1)
            if (k <= max_cluster && tiles * k <= num_sms) { csize = k; break; }
        static const int force_cluster = getenv("ETP_CLUSTER_FORCE") ? atoi(getenv("ETP_CLUSTER_FORCE")) : 0;  // experiments
        if (force_cluster > 1 && !cluster_refused) csize = force_cluster;
        if (csize > 1) {
            BatchArgs b2 = *ba;
            b2.dbg_min_len = dbg_min_len;
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3((unsigned)(tiles * csize));
            cfg.blockDim = dim3(kThreads);
            cfg.dynamicSmemBytes = smem;
            cfg.stream = stream;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = (unsigned)csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at;
            cfg.numAttrs = 1;
            const cudaError_t ce = cudaLaunchKernelEx(&cfg, score_kernel<R, NT, kSym, kGiven, kBatch>, st, pr, out, counters, b2);
            if (ce == cudaSuccess) {
                *grid_out = (int)(tiles * csize);
                return ce;
            }
            // the first cycle of a shape is always launched outside a capture, so a device (partition) that cannot co-schedule
            // the cluster is found out here: one CTA per tile instead
            cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
            cudaStreamIsCapturing(stream, &cap);
            if (cap != cudaStreamCaptureStatusNone) return ce;
            (void)cudaGetLastError();
            cluster_refused = true;
        }
    }
    score_kernel<R, NT, kSym, kGiven, kBatch><<<(int)grid, kThreads, smem, stream>>>(st, pr, out, counters, kBatch ? *ba : BatchArgs{});
    return cudaGetLastError();
}

cudaError_t launch_score(int precision, const DevStep& st, const DevParams& pr, const DevOut& out, unsigned* counters, int num_sms,
                         int* grid_out, cudaStream_t stream, const BatchArgs* ba, size_t batch_smem) {
#define ETP_LAUNCH(R, NT, SYM, GIVEN) \
    launch_score_t<R, NT, SYM, GIVEN>(st, pr, out, counters, num_sms, grid_out, ba, batch_smem, stream)
    if (ba) {  // scenario sweep: the default model's specialisation, or the generic tables
        if (precision == ETP_PRECISION_FP64) return launch_score_t<double, -1, false, false, true>(st, pr, out, counters, num_sms, grid_out, ba, batch_smem, stream);
        bool lr4s = pr.n_thr == 2 && pr.f_fcos[0] == -pr.f_fcos[1];
        for (int i = 0; i <= pr.n_thr; ++i) lr4s = lr4s && pr.coef_pos[i] == pr.coef_neg[i];
        if (lr4s) return launch_score_t<float, 2, true, false, true>(st, pr, out, counters, num_sms, grid_out, ba, batch_smem, stream);
        return launch_score_t<float, -1, false, false, true>(st, pr, out, counters, num_sms, grid_out, ba, batch_smem, stream);
    }
    if (st.given) {  // caller-provided trajectories: the API-compatibility path, generic tables only
        if (precision == ETP_PRECISION_FP64) return ETP_LAUNCH(double, -1, false, true);
        return ETP_LAUNCH(float, -1, false, true);
    }
    if (precision == ETP_PRECISION_FP64) return ETP_LAUNCH(double, -1, false, false);
    bool sym = true;  // LR4S / LR12S: one coefficient table for both sides
    for (int i = 0; i <= pr.n_thr; ++i) sym = sym && pr.coef_pos[i] == pr.coef_neg[i];
    bool paired = true;  // thresholds symmetric about pi/2 (chain_coef)
    for (int i = 0; i < pr.n_thr; ++i) paired = paired && pr.f_fcos[i] == -pr.f_fcos[pr.n_thr - 1 - i];
    switch (paired ? pr.n_thr : -1) {  // the reference's models: LR1S (0 thresholds), LR4S / LR4A (2), LR12S / LR12A (6)
        case 0: return ETP_LAUNCH(float, 0, true, false);
        case 2: return sym ? ETP_LAUNCH(float, 2, true, false) : ETP_LAUNCH(float, 2, false, false);
        case 6: return sym ? ETP_LAUNCH(float, 6, true, false) : ETP_LAUNCH(float, 6, false, false);
        default: return ETP_LAUNCH(float, -1, false, false);
    }
#undef ETP_LAUNCH
}

size_t score_smem_bytes(int precision, int Tmax, int O) {
    return precision == ETP_PRECISION_FP64 ? Smem<double>::bytes(Tmax, O) : Smem<float>::bytes(Tmax, O);
}

cudaError_t read_debug_counters(unsigned long long* out, int n) {
    unsigned long long v = 0;
    cudaError_t e = cudaMemcpyFromSymbol(&v, g_fallback_evals, sizeof(v));
    if (n > 0) out[0] = v;
    return e;
}

}  // namespace etp
