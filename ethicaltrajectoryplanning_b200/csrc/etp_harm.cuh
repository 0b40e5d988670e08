// Typed helpers and the harm model shared by the fused scoring kernel (etp_score.cu) and the exact re-scoring of
// near-tie candidates (etp_select.cu).  Citations are to the reference tree, paths relative to its root.
#pragma once

#include "etp_bvn.cuh"
#include "etp_device.cuh"

namespace etp {

constexpr unsigned kFull = 0xffffffffu;

// ---------------------------------------------------------------------------------------------
// small typed helpers
// ---------------------------------------------------------------------------------------------
// order-preserving integer image of a real: max over reals == max over images (shared-memory atomicMax)
template <typename R> struct Enc;
template <> struct Enc<float> {
    using U = unsigned int;
    static __device__ __forceinline__ U enc(float x) {
        U b = __float_as_uint(x);
        return (b >> 31) ? ~b : (b | 0x80000000u);
    }
    static __device__ __forceinline__ float dec(U k) {
        U b = (k >> 31) ? (k & 0x7fffffffu) : ~k;
        return __uint_as_float(b);
    }
};
template <> struct Enc<double> {
    using U = unsigned long long;
    static __device__ __forceinline__ U enc(double x) {
        U b = (U)__double_as_longlong(x);
        return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
    }
    static __device__ __forceinline__ double dec(U k) {
        U b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
        return __longlong_as_double((long long)b);
    }
};

template <typename R> __device__ __forceinline__ Vec4<R> ldg4(const Vec4<R>* p);
template <> __device__ __forceinline__ Vec4<float> ldg4<float>(const Vec4<float>* p) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    return Vec4<float>{v.x, v.y, v.z, v.w};
}
template <> __device__ __forceinline__ Vec4<double> ldg4<double>(const Vec4<double>* p) {
    const double2 a = __ldg(reinterpret_cast<const double2*>(p)), b = __ldg(reinterpret_cast<const double2*>(p) + 1);
    return Vec4<double>{a.x, a.y, b.x, b.y};
}
template <typename R> __device__ __forceinline__ R rmax_(R a, R b) { return a > b ? a : b; }
// square root of the fp32 streaming path: MUFU.RSQ based (2 ulp); the fp64 check mode stays IEEE
__device__ __forceinline__ float sqrt_fast(float x) {
    const float xp = fmaxf(x, 0.0f);  // rounding can leave (v_e - v_o)^2 slightly negative
    return xp * rsqrt_ftz(fmaxf(xp, 1e-30f));
}
__device__ __forceinline__ double sqrt_fast(double x) { return sqrt(x); }
__device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
__device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
__device__ __forceinline__ float exp_(float x) { return expf(x); }
__device__ __forceinline__ double exp_(double x) { return exp(x); }
__device__ __forceinline__ float abs_(float x) { return fabsf(x); }
__device__ __forceinline__ double abs_(double x) { return fabs(x); }

// Constants of the risk model in the arithmetic type R, resident in shared memory (one copy per CTA):
// dynamic indexing into kernel parameters would force a per-thread local-memory copy of DevParams.
template <typename R> struct Tab {
    R thr[ETP_MAX_ANGLE_BINS], fcos[ETP_MAX_ANGLE_BINS], coef_pos[ETP_MAX_ANGLE_BINS + 1], coef_neg[ETP_MAX_ANGLE_BINS + 1];
    R lr_speed, lr1s_speed, ped_speed, lr_const, lr1s_const, neg_ped_const;
    R eps_pos;     // fp32 rounding of a position difference in metres (2 ulp of the largest coordinate in play)
    R hx, hy, rr;  // ego box half extents (l/6, w/2) and centre offset (l/2)(2/3)  (collision_probability.py:157,349-362)
    int n_thr;
};

template <typename R> __device__ __forceinline__ void fill_tab(Tab<R>& tb, const DevParams& pr, double eps_pos) {
    for (int i = 0; i < ETP_MAX_ANGLE_BINS; ++i) { tb.thr[i] = (R)pr.thr[i]; tb.fcos[i] = (R)pr.fcos[i]; }
    for (int i = 0; i <= ETP_MAX_ANGLE_BINS; ++i) { tb.coef_pos[i] = (R)pr.coef_pos[i]; tb.coef_neg[i] = (R)pr.coef_neg[i]; }
    tb.lr_speed = (R)pr.lr_speed; tb.lr1s_speed = (R)pr.lr1s_speed; tb.ped_speed = (R)pr.ped_speed;
    tb.lr_const = (R)pr.lr_const; tb.lr1s_const = (R)pr.lr1s_const; tb.neg_ped_const = (R)(-pr.ped_const);
    tb.hx = (R)(pr.veh_l / 6); tb.hy = (R)(pr.veh_w / 2); tb.rr = (R)((pr.veh_l / 2) * (2.0 / 3.0));
    tb.n_thr = pr.n_thr;
    tb.eps_pos = (R)eps_pos;
}

// ---------------------------------------------------------------------------------------------
// harm model (harm_estimation.py:313-332): harm = 1 / (1 + exp(-(const + arg))), arg = speed * dv_k + coef(angle)
// ---------------------------------------------------------------------------------------------
// delta_v = sqrt(v_e^2 + v_o^2 + 2 v_e v_o cos(pdof)), pdof = psi_o - psi_e + pi (harm_estimation.py:313,322-325)
__device__ __forceinline__ float delta_v(float ec, float es, float ev, float, float ovo, float ocp, float osp, float) {
    const float cpd = -(ocp * ec + osp * es);
    return sqrt_fast(ev * ev + ovo * ovo + 2 * ev * ovo * cpd);
}
// fp64 check mode evaluates pdof with the reference's own expression
__device__ __forceinline__ double delta_v(double, double, double ev, double eyaw, double ovo, double, double, double opsi) {
    const double pdof = opsi - eyaw + 3.14159265358979323846;
    return sqrt(ev * ev + ovo * ovo + 2 * ev * ovo * cos(pdof));
}

// impact-area coefficient (logistic_regression_*.py): thresholds on |angle|, sign picks the side
template <typename R> __device__ __forceinline__ R angle_coef(const Tab<R>& pr, R angle) {
    const R a = abs_(angle);
    int i = 0;
    while (i < pr.n_thr && !(a < pr.thr[i])) ++i;
    return angle < (R)0 ? pr.coef_neg[i] : pr.coef_pos[i];
}

// Band of an impact angle given as a direction: P = |d| cos(angle), d2 = |d|^2.  |angle| < thr is
// cos(angle) > cos(thr), compared through the sign-preserving squares P|P| > d2 * cos(thr)|cos(thr)|.
// `tol` bounds what fp32 rounding (of the positions above all) can move a test by; a test closer to its threshold
// than that sets `amb`, and the caller repeats the decision in fp64 from the exact inputs (band_coefs_exact).
__device__ __forceinline__ int direction_band(const Tab<float>& pr, float P, float d2, float tol, bool& amb) {
    // thresholds increase, so the tests pass from some band on: the band is the number of failed tests
    const float fp = P * fabsf(P);
    int band = 0;
    for (int i = 0; i < pr.n_thr; ++i) {
        const float df = fmaf(-d2, pr.fcos[i], fp);
        band += (df > 0.0f) ? 0 : 1;
        amb = amb || fabsf(df) < tol;
    }
    return band;
}

// tolerance of a band test: the direction (dx, dy) carries the rounding of two positions (eps_pos metres in total),
// which moves P|P| - d2 f by at most |d| eps_pos <= sqrt(2) max(|dx|, |dy|) eps_pos; the arithmetic adds a few ulps of d2
__device__ __forceinline__ float band_tol(float d2, float dx, float dy, float eps_pos) {
    return __fmaf_rn(fmaxf(fabsf(dx), fabsf(dy)), __fmul_rn(1.4143f, eps_pos), __fmul_rn(d2, 1e-6f));
}

// Impact-area coefficients of ego and obstacle, fp32 path.  `ewrap` are the whole turns that bring the ego yaw
// into (-pi, pi].  The reference bins the UN-wrapped angles rel - yaw_e and
// pi + rel - psi_o (quirk Q4) with rel = atan2(dy, dx); instead of evaluating atan2, the relative direction is
// rotated into the ego / obstacle frame (P, Q) and the number of 2 pi wraps of the un-wrapped difference is
// recovered from the half planes of the two directions.  |angle| > pi always lands in the last ("rear") band.
__device__ __forceinline__ void impact_coefs(const Tab<float>& pr, float ex, float ey, float ec, float es, float ewrap,
                                             float mx, float my, float ocp, float osp, float ocode, float& ce, float& co,
                                             bool& amb) {
    const float dx = mx - ex, dy = my - ey;
    const float d2 = dx * dx + dy * dy;
    const float tol = band_tol(d2, dx, dy, pr.eps_pos);
    const int last = pr.n_thr;
    // ego: angle = rel - yaw_e
    const float p = dx * ec + dy * es, q = dy * ec - dx * es;
    int n0 = 0;
    if (dy >= 0 && es < 0 && q < 0) n0 = 1;
    if (dy < 0 && es >= 0 && q > 0) n0 = -1;
    const int be = ((float)n0 != ewrap) ? last : direction_band(pr, p, d2, tol, amb);
    ce = q < 0 ? pr.coef_neg[be] : pr.coef_pos[be];
    // obstacle: angle = pi + rel - psi_o; `ocode` is the validity table of pack_obstacles_kernel
    const float po = dx * ocp + dy * osp, qo = dy * ocp - dx * osp;
    const bool pos = qo < 0 || (qo == 0 && po >= 0);
    const bool valid = ((__float_as_int(ocode) >> ((pos ? 2 : 0) | (dy >= 0 ? 1 : 0))) & 1) != 0;
    const bool neg_side = !pos && valid;
    const int bo = valid ? direction_band(pr, -po, d2, tol, amb) : last;
    co = neg_side ? pr.coef_neg[bo] : pr.coef_pos[bo];
}
__device__ __forceinline__ void impact_coefs(const Tab<double>& pr, double ex, double ey, double, double, double eyaw,
                                             double mx, double my, double, double, double opsi, double& ce, double& co,
                                             bool&) {
    const double rel = atan2(my - ey, mx - ex);  // harm_estimation.py:314-319
    ce = angle_coef<double>(pr, rel - eyaw);
    co = angle_coef<double>(pr, 3.14159265358979323846 + rel - opsi);
}

template <typename R> __device__ __forceinline__ R sigmoid_harm(R cst, R arg) { return (R)1 / ((R)1 + exp_(-cst - arg)); }
// the same with two MUFU operations (ex2, rcp): ~6e-7 relative, for the harm that only multiplies a probability
__device__ __forceinline__ float sigmoid_risk(float cst, float arg) {
    return rcp_ftz(1.0f + ex2_ftz(-1.4426950408889634f * (cst + arg)));
}
__device__ __forceinline__ double sigmoid_risk(double cst, double arg) { return sigmoid_harm<double>(cst, arg); }

// 1e-4 / mahalanobis(ego[i], pos[i-1], inv(cov[i-1]))  (collision_probability.py:281-289)
template <typename R> __device__ __forceinline__ R inv_mahalanobis(R x, R y, R mx, R my, R i00, R i01, R i11) {
    const R dx = x - mx, dy = y - my;
    const R m2 = dx * dx * i00 + dx * dy * i01 + dy * dy * i11;
    return (R)1e-4 / sqrt_(m2);
}

}  // namespace etp
