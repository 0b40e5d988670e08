// Scoring kernels of the B200-native candidate-trajectory path (sm_100a).
//
//   pack_obstacles_kernel : predictions (SoA fp64) -> packed obstacle tile [field][o][t] of R
//   profile_kernel        : per (v,t) profile stage in fp64: quartic, spline lookup, chained
//                           non-uniform gradients, reference yaw / curvature / curvature'
//                           (frenet_functions.py:258-325)
//   score_kernel<R>       : fused per-candidate pass: quintic + Frenet->Cartesian + validity masks,
//                           collision probability, harm, risk maxima over time, ethical principle
//                           costs, weighted cost, validity level, block arg-min; last block
//                           finishes the grid arg-min (frenet_functions.py:330-474, 529-593;
//                           collision_probability.py:136-256; harm_estimation.py:306-332;
//                           risk_costs.py:87-101,128-255; calc_trajectory_cost.py:103-146,279-346)
//   trajectory_kernel     : one candidate recomputed in fp64 for the host (frenet_planner.py:564-576)
#include <cooperative_groups.h>

#include "etp_bvn.cuh"
#include "etp_device.cuh"
#include "etp_launch.h"

namespace etp {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kQueue = 64;  // per-warp deferred (un-gated) evaluations
constexpr unsigned kFull = 0xffffffffu;

// ---------------------------------------------------------------------------------------------
// small typed helpers
// ---------------------------------------------------------------------------------------------
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

template <typename R> struct alignas(16) Vec4 { R x, y, z, w; };
template <typename R> struct alignas(2 * sizeof(R)) Vec2 { R x, y; };
// read-only 16-byte loads of the obstacle vector planes
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
__device__ __forceinline__ float sqrt_fast(float x) { return x > 0.0f ? x * rsqrtf(x) : 0.0f; }
__device__ __forceinline__ double sqrt_fast(double x) { return sqrt(x); }
__device__ __forceinline__ float rsqrt_(float x) { return sqrtf(x); }
__device__ __forceinline__ float sqrt_(float x) { return sqrtf(x); }
__device__ __forceinline__ double sqrt_(double x) { return sqrt(x); }
__device__ __forceinline__ float atan2_(float y, float x) { return atan2f(y, x); }
__device__ __forceinline__ double atan2_(double y, double x) { return atan2(y, x); }
__device__ __forceinline__ float cos_(float x) { return cosf(x); }
__device__ __forceinline__ double cos_(double x) { return cos(x); }
__device__ __forceinline__ float exp_(float x) { return expf(x); }
__device__ __forceinline__ double exp_(double x) { return exp(x); }
__device__ __forceinline__ float abs_(float x) { return fabsf(x); }
__device__ __forceinline__ double abs_(double x) { return fabs(x); }

template <typename T> __device__ __forceinline__ T warp_max(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        T u = __shfl_xor_sync(kFull, v, o);
        v = u > v ? u : v;
    }
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------
// pack_obstacles_kernel
// ---------------------------------------------------------------------------------------------
template <typename R>
__global__ void pack_obstacles_kernel(int O, int Tp, const int* __restrict__ pred_len, const double* __restrict__ pos,
                                      const double* __restrict__ cov, const double* __restrict__ yaw,
                                      const double* __restrict__ vel, const double* __restrict__ length,
                                      const double* __restrict__ mass, const int* __restrict__ prot,
                                      const int* __restrict__ resp, double veh_m, R* __restrict__ tile,
                                      double* __restrict__ oconst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= O * Tp) return;
    const int o = i / Tp, t = i % Tp;
    const int n = pred_len[o];
    const size_t plane = (size_t)O * Tp;
    // tile = [V0: (mx,my,ex,ey)][V1: (vo,cos psi,sin psi,angle)] as Vec4 planes [t][o], then the scalar planes [field][t][o]
    Vec4<R>* v0 = reinterpret_cast<Vec4<R>*>(tile);
    Vec4<R>* v1 = v0 + plane;
    R* out = reinterpret_cast<R*>(v1 + plane) + (size_t)t * O + o;
    double v[OF_COUNT];
    for (int f = 0; f < OF_COUNT; ++f) v[f] = 0.0;
    if (t < n) {
        const double* p = pos + ((size_t)o * Tp + t) * 2;
        const double* c = cov + ((size_t)o * Tp + t) * 4;
        v[OF_MX] = p[0];
        v[OF_MY] = p[1];
        if (t + 1 < n) {
            const double y1 = yaw[(size_t)o * Tp + t + 1];
            // np.stack((cos, sin)) * length / 2   (collision_probability.py:175)
            v[OF_EX] = cos(y1) * length[o] / 2;
            v[OF_EY] = sin(y1) * length[o] / 2;
        }
        double c00 = c[0], c01 = c[1], c10 = c[2], c11 = c[3];
        if (c00 == 0.0 && c01 == 0.0 && c10 == 0.0 && c11 == 0.0) {  // collision_probability.py:210-212
            c00 = 0.1; c11 = 0.1;
        }
        const double sx = sqrt(c00), sy = sqrt(c11);
        v[OF_ISX] = 1.0 / sx;
        v[OF_ISY] = 1.0 / sy;
        v[OF_RHO] = c10 / sy / sx;  // mvnun takes the correlation from the lower triangle
        const double psi = yaw[(size_t)o * Tp + t];
        v[OF_PSI] = psi;
        v[OF_VO] = vel[(size_t)o * Tp + t];
        v[OF_CPSI] = cos(psi);
        v[OF_SPSI] = sin(psi);
        v[OF_WRAP] = ceil((psi - 3.14159265358979323846) / 6.28318530717958647692);
        // Mahalanobis mode inverts cov_list itself (collision_probability.py:279), not the 0.1*I stand-in
        const double r00 = c[0], r01 = c[1], r10 = c[2], r11 = c[3];
        const double det = r00 * r11 - r01 * r10;
        v[OF_I00] = r11 / det;
        v[OF_I01] = -(r01 + r10) / det;
        v[OF_I11] = r00 / det;
    }
    for (int f = 0; f < OF_COUNT; ++f) out[f * plane] = (R)v[f];
    const size_t vi = (size_t)t * O + o;
    v0[vi] = Vec4<R>{(R)v[OF_MX], (R)v[OF_MY], (R)v[OF_EX], (R)v[OF_EY]};
    v1[vi] = Vec4<R>{(R)v[OF_VO], (R)v[OF_CPSI], (R)v[OF_SPSI], (R)v[sizeof(R) == 4 ? OF_WRAP : OF_PSI]};
    if (t == 0) {
        const double mo = mass[o];
        oconst[o * OC_COUNT + OC_KE] = mo / (veh_m + mo);     // harm_estimation.py:327
        oconst[o * OC_COUNT + OC_KO] = veh_m / (veh_m + mo);  // harm_estimation.py:328
        oconst[o * OC_COUNT + OC_PROT] = prot[o] ? 1.0 : 0.0;
        oconst[o * OC_COUNT + OC_RESP] = (double)resp[o];
        oconst[o * OC_COUNT + OC_HALFLEN] = length[o] / 2;
    }
}

// ---------------------------------------------------------------------------------------------
// profile_kernel: one block per (v, t) profile, everything fp64
// ---------------------------------------------------------------------------------------------
// 32-ary search of the spline segment with warp ballots: the whole warp locates one abscissa.
__device__ __forceinline__ int warp_find_segment(const double* __restrict__ knots, int nseg, double s, int lane) {
    // largest i in [0, nseg-1] with knots[i] <= s (clamped): searchsorted(side='right') - 1
    int lo = 0, hi = nseg;  // candidates lo .. hi-1
    while (hi - lo > 1) {
        const int span = hi - lo;
        const int step = (span + 31) / 32;
        const int idx = lo + lane * step;
        const bool le = (idx < hi) && (knots[idx] <= s);
        const unsigned m = __ballot_sync(kFull, le);
        // lanes are monotone: knots increasing -> predicate is a prefix of ones (possibly empty)
        const int ones = __popc(m);
        if (ones == 0) return lo;  // s below the first knot of the window: clamp
        const int base = lo + (ones - 1) * step;
        lo = base;
        hi = min(base + step, hi);
    }
    return lo;
}

// numpy.gradient(f, s) for non-uniform spacing: second order inside, first order at the ends
__device__ __forceinline__ double np_gradient(const double* f, const double* s, int i, int n) {
    if (i == 0) return (f[1] - f[0]) / (s[1] - s[0]);
    if (i == n - 1) return (f[n - 1] - f[n - 2]) / (s[n - 1] - s[n - 2]);
    const double dx1 = s[i] - s[i - 1], dx2 = s[i + 1] - s[i];
    const double a = -(dx2) / (dx1 * (dx1 + dx2));
    const double b = (dx2 - dx1) / (dx1 * dx2);
    const double c = dx1 / (dx2 * (dx1 + dx2));
    return a * f[i - 1] + b * f[i] + c * f[i + 1];
}

__global__ void __launch_bounds__(128) profile_kernel(DevStep st, DevParams pr) {
    extern __shared__ double sm[];
    const int Tmax = st.Tmax;
    double* s_s = sm;
    double* s_x = sm + Tmax;
    double* s_y = sm + 2 * Tmax;
    double* s_a = sm + 3 * Tmax;  // dx   -> dddx
    double* s_b = sm + 4 * Tmax;  // dy   -> dddy
    double* s_c = sm + 5 * Tmax;  // ddx
    double* s_d = sm + 6 * Tmax;  // ddy
    const int p = st.p_begin + blockIdx.x;
    const int iv = p / st.nt, it = p % st.nt;
    const double vT = st.v_list[iv], tT = st.t_list[it];
    const int T = st.nsamp[it];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* pf = st.prof + (size_t)blockIdx.x * PF_COUNT * Tmax;

    // quartic_polynomial(xs=c_s, vxs=c_s_d, axs=c_s_dd, vxe=vT, axe=0, T=tT-dt)  (frenet_functions.py:258-260)
    const double Te = tT - st.dt;
    const double a0 = st.c_s, a1 = st.c_s_d, a2 = st.c_s_dd / 2.0;
    const double b1 = vT - a1 - 2 * a2 * Te, b2 = 0.0 - 2 * a2;
    const double a3 = (3.0 * b1 - b2 * Te) / (3.0 * Te * Te);
    const double a4 = (b2 * Te - 2.0 * b1) / (4.0 * Te * Te * Te);

    int vel_ok = 1;
    for (int t = tid; t < T; t += blockDim.x) {
        const double tt = t * st.dt;
        const double t2 = tt * tt, t3 = t2 * tt, t4 = t2 * t2;
        const double s = a0 + a1 * tt + a2 * t2 + a3 * t3 + a4 * t4;
        const double sd = a1 + 2 * a2 * tt + 3 * a3 * t2 + 4 * a4 * t3;
        s_s[t] = s;
        pf[PF_S * Tmax + t] = s;
        pf[PF_S_D * Tmax + t] = sd;
        pf[PF_S_DD * Tmax + t] = 2 * a2 + 6 * a3 * tt + 12 * a4 * t2;
        pf[PF_S_DDD * Tmax + t] = 6 * a3 + 24 * a4 * tt;
        if (!(fabs(sd) <= pr.v_max)) vel_ok = 0;  // validity_checks.py:158
    }
    vel_ok = __syncthreads_and(vel_ok);
    // reference position: one warp per abscissa, 32-ary segment search (csp.calc_position, :278)
    for (int t = warp; t < T; t += blockDim.x / 32) {
        const double s = s_s[t];
        const int seg = warp_find_segment(st.knots, st.nseg, s, lane);
        if (lane == 0) {
            const double ds = s - st.knots[seg];
            const double* cx = st.cx + 4 * seg;
            const double* cy = st.cy + 4 * seg;
            s_x[t] = cx[0] + cx[1] * ds + cx[2] * ds * ds + cx[3] * ds * ds * ds;
            s_y[t] = cy[0] + cy[1] * ds + cy[2] * ds * ds + cy[3] * ds * ds * ds;
        }
    }
    __syncthreads();
    for (int t = tid; t < T; t += blockDim.x) {  // dx, dy (:291,294)
        s_a[t] = np_gradient(s_x, s_s, t, T);
        s_b[t] = np_gradient(s_y, s_s, t, T);
    }
    __syncthreads();
    for (int t = tid; t < T; t += blockDim.x) {  // ddx, ddy (:292,295)
        s_c[t] = np_gradient(s_a, s_s, t, T);
        s_d[t] = np_gradient(s_b, s_s, t, T);
    }
    __syncthreads();
    for (int t = tid; t < T; t += blockDim.x) {
        const double dx = s_a[t], dy = s_b[t], ddx = s_c[t], ddy = s_d[t];
        const double dddx = np_gradient(s_c, s_s, t, T), dddy = np_gradient(s_d, s_s, t, T);  // :293,296
        const double yaw = atan2(dy, dx);  // :302
        // :308-310 -- the denominator is dx^2 + (dy^2)^(3/2) in the reference (operator precedence)
        const double curv = (dx * ddy - ddx * dy) / (dx * dx + pow(dy * dy, 1.5));
        // :316-325
        const double z = dx * ddy - ddx * dy;
        const double z_d = dx * dddy - dddx * dy;
        const double q = dx * dx + dy * dy;
        const double n = pow(q, 1.5);
        const double n_d = 1.5 * (sqrt(q) * (2 * (dx * ddx) + 2 * (dy * ddy)));
        const double curv_d = (z_d * n - z * n_d) / (n * n);
        pf[PF_GX * Tmax + t] = s_x[t];
        pf[PF_GY * Tmax + t] = s_y[t];
        pf[PF_YAW * Tmax + t] = yaw;
        pf[PF_CURV * Tmax + t] = curv;
        pf[PF_CURV_D * Tmax + t] = curv_d;
        pf[PF_SIN * Tmax + t] = sin(yaw);
        pf[PF_COS * Tmax + t] = cos(yaw);
    }
    for (int t = T + tid; t < Tmax; t += blockDim.x)
        for (int f = 0; f < PF_COUNT; ++f) pf[f * Tmax + t] = 0.0;
    if (tid == 0) {
        double* pm = st.prof_meta + (size_t)blockIdx.x * PM_COUNT;
        pm[PM_DS] = s_s[T - 1] - s_s[0];                                                  // :327-328
        pm[PM_LOW] = (fabs(st.c_s_d) < st.v_thr || fabs(vT) < st.v_thr) ? 1.0 : 0.0;  // :248-251
        pm[PM_T] = (double)T;
        pm[PM_VEL_OK] = (double)vel_ok;
        pm[PM_TEND] = Te;
    }
}

// ---------------------------------------------------------------------------------------------
// score_kernel: persistent warps, one candidate per warp at a time
// ---------------------------------------------------------------------------------------------
// Work decomposition.  A candidate (one lateral target of one (v,t) profile) is owned by ONE warp for
// its whole life: generation and Frenet->Cartesian transform (lanes over time), risk rows
// (obstacle by obstacle, lanes over time, reductions over time with warp shuffles), principle
// costs (lanes over obstacles), level, cost and arg-min key.  Warps pull chunks of consecutive
// candidates from a global counter, so there is no block barrier inside the loop and no static
// imbalance between obstacles or profiles.  Evaluations that pass the 5 m gate are deferred to a
// warp-private queue and executed 32 at a time so that the expensive path runs with full warps.
enum { EG_X = 0, EG_Y, EG_C, EG_S, EG_V, EG_YAW, EG_WRAP, EG_COUNT };
constexpr int kChunk = 4;  // consecutive candidates per work grab (same profile rows stay in L1)

template <typename R> __host__ __device__ inline size_t warp_smem_bytes(int Ts, int O) {
    const int On = O > 0 ? O : 1;
    size_t b = sizeof(R) * 4 * Ts;                      // ego (x, y, cos yaw, sin yaw)
    b += sizeof(double) * 2 * Ts;                       // exact ego x, y
    b += sizeof(double) * 2 * On;                       // risk maxima (ordered-int encoded), sized for fp64
    b += sizeof(R) * 2 * Ts;                            // ego (v, yaw angle or its wrap count)
    b += sizeof(R) * 2 * On;                            // harm logit maxima
    b += sizeof(unsigned) * kQueue;                     // deferred evaluations
    return (b + 15) & ~size_t(15);
}

template <typename R> struct WarpSmem {
    Vec4<R>* e0;               // [Ts] x, y, cos yaw, sin yaw
    double* xd;
    double* yd;
    typename Enc<R>::U* rmax;  // [2][On]
    Vec2<R>* e1;               // [Ts] v, angle (fp32: wrap count of the yaw; fp64: the yaw)
    R* hmax;                   // [2][On]
    unsigned* queue;
    __device__ __forceinline__ WarpSmem(unsigned char* base, int Ts, int O) {
        const int On = O > 0 ? O : 1;
        e0 = reinterpret_cast<Vec4<R>*>(base);
        double* d = reinterpret_cast<double*>(e0 + Ts);
        xd = d; d += Ts;
        yd = d; d += Ts;
        rmax = reinterpret_cast<typename Enc<R>::U*>(d); d += 2 * On;
        e1 = reinterpret_cast<Vec2<R>*>(d);
        hmax = reinterpret_cast<R*>(e1 + Ts);
        queue = reinterpret_cast<unsigned*>(hmax + 2 * On);
    }
};

template <typename R> struct Obs {
    const Vec4<R>* v0;  // (mx, my, ex, ey)            [t][o]
    const Vec4<R>* v1;  // (vo, cos psi, sin psi, angle) [t][o]
    const R* base;      // scalar planes [field][t][o]
    int plane;
    int O;
    __device__ __forceinline__ Obs(const void* tile, int O_, int Tp_) {
        plane = O_ * Tp_;
        O = O_;
        v0 = reinterpret_cast<const Vec4<R>*>(tile);
        v1 = v0 + plane;
        base = reinterpret_cast<const R*>(v1 + plane);
    }
    __device__ __forceinline__ R get(int f, int o, int t) const { return __ldg(base + (f * plane + t * O + o)); }
};

// Constants of the risk model in the arithmetic type R, resident in shared memory (one copy per CTA):
// dynamic indexing into kernel parameters would force a per-thread local-memory copy of DevParams.
template <typename R> struct Tab {
    R thr[ETP_MAX_ANGLE_BINS], fcos[ETP_MAX_ANGLE_BINS], coef_pos[ETP_MAX_ANGLE_BINS + 1], coef_neg[ETP_MAX_ANGLE_BINS + 1];
    R lr_speed, lr1s_speed, ped_speed, lr_const, lr1s_const, neg_ped_const;
    R hx, hy, rr;  // ego box half extents (l/6, w/2) and centre offset (l/2)(2/3)  (collision_probability.py:157,349-362)
    int n_thr;
};

template <typename R> __device__ __forceinline__ void fill_tab(Tab<R>& tb, const DevParams& pr) {
    for (int i = 0; i < ETP_MAX_ANGLE_BINS; ++i) { tb.thr[i] = (R)pr.thr[i]; tb.fcos[i] = (R)pr.fcos[i]; }
    for (int i = 0; i <= ETP_MAX_ANGLE_BINS; ++i) { tb.coef_pos[i] = (R)pr.coef_pos[i]; tb.coef_neg[i] = (R)pr.coef_neg[i]; }
    tb.lr_speed = (R)pr.lr_speed; tb.lr1s_speed = (R)pr.lr1s_speed; tb.ped_speed = (R)pr.ped_speed;
    tb.lr_const = (R)pr.lr_const; tb.lr1s_const = (R)pr.lr1s_const; tb.neg_ped_const = (R)(-pr.ped_const);
    tb.hx = (R)(pr.veh_l / 6); tb.hy = (R)(pr.veh_w / 2); tb.rr = (R)((pr.veh_l / 2) * (2.0 / 3.0));
    tb.n_thr = pr.n_thr;
}

// impact-area coefficient (logistic_regression_*.py): thresholds on |angle|, sign picks the side
template <typename R> __device__ __forceinline__ R angle_coef(const Tab<R>& pr, R angle) {
    const R a = abs_(angle);
    int i = 0;
    while (i < pr.n_thr && !(a < pr.thr[i])) ++i;
    return angle < (R)0 ? pr.coef_neg[i] : pr.coef_pos[i];
}

// harm "logit arguments" of ego and obstacle at trajectory index t against obstacle sample t
// (harm_estimation.py:313-332); harm = 1 / (1 + exp(-(const + arg))).
// Band of an impact angle given as a direction: P = |d| cos(angle), d2 = |d|^2.  |angle| < thr is
// cos(angle) > cos(thr), compared through the sign-preserving squares P|P| > d2 * cos(thr)|cos(thr)|.
__device__ __forceinline__ int direction_band(const Tab<float>& pr, float P, float d2) {
    // thresholds increase, so the tests pass from some band on: the band is the number of failed tests
    const float fp = P * fabsf(P);
    int band = 0;
    for (int i = 0; i < pr.n_thr; ++i) band += (fp > d2 * pr.fcos[i]) ? 0 : 1;
    return band;
}

// fp32 path.  `ewrap` / `owrap` are the whole turns that bring the ego / obstacle yaw into (-pi, pi].
// The reference bins the UN-wrapped angles rel - yaw_e and pi + rel - psi_o (quirk Q4) with
// rel = atan2(dy, dx); instead of evaluating atan2, the relative direction is rotated into the ego /
// obstacle frame (P, Q) and the number of 2 pi wraps of the un-wrapped difference is recovered from
// the half planes of the two directions.  |angle| > pi always lands in the last ("rear") band.
template <typename R>
__device__ __forceinline__ void harm_args(const Tab<R>& pr, bool prot, R ke, R ko, R ex, R ey, R ec, R es, R ev, R ewrap,
                                          R mx, R my, R owrap, R ovo, R ocp, R osp, R& arg_e, R& arg_o) {
    // delta_v = sqrt(v_e^2 + v_o^2 + 2 v_e v_o cos(pdof)), pdof = psi_o - psi_e + pi
    const R cpd = -(ocp * ec + osp * es);
    const R dv = sqrt_fast(ev * ev + ovo * ovo + 2 * ev * ovo * cpd);
    const R dve = ke * dv, dvo = ko * dv;
    if (prot) {
        const R dx = mx - ex, dy = my - ey;
        const R d2 = dx * dx + dy * dy;
        const int last = pr.n_thr;
        // ego: angle = rel - yaw_e
        const R p = dx * ec + dy * es, q = dy * ec - dx * es;
        int n0 = 0;
        if (dy >= 0 && es < 0 && q < 0) n0 = 1;
        if (dy < 0 && es >= 0 && q > 0) n0 = -1;
        const int be = ((R)n0 != ewrap) ? last : direction_band(pr, p, d2);
        const R ce = q < 0 ? pr.coef_neg[be] : pr.coef_pos[be];
        // obstacle: angle = pi + rel - psi_o
        const R po = dx * ocp + dy * osp, qo = dy * ocp - dx * osp;
        n0 = 0;
        if (dy >= 0 && osp < 0 && qo < 0) n0 = 1;
        if (dy < 0 && osp >= 0 && qo > 0) n0 = -1;
        const R k = (R)n0 - owrap;
        const bool pos_side = (k == 0) && (qo <= 0) && !(qo == 0 && po < 0);
        const bool neg_side = (k == (R)-1) && (qo >= 0);
        const int bo = (pos_side || neg_side) ? direction_band(pr, -po, d2) : last;
        const R co = neg_side ? pr.coef_neg[bo] : pr.coef_pos[bo];
        arg_e = pr.lr_speed * dve + ce;
        arg_o = pr.lr_speed * dvo + co;
    } else {
        arg_e = pr.lr1s_speed * dve;
        arg_o = pr.ped_speed * dvo;
    }
}
// fp64 check mode evaluates pdof with the reference's own expression
template <>
__device__ __forceinline__ void harm_args<double>(const Tab<double>& pr, bool prot, double ke, double ko, double ex, double ey,
                                                  double ec, double es, double ev, double eyaw, double mx, double my,
                                                  double opsi, double ovo, double ocp, double osp, double& arg_e,
                                                  double& arg_o) {
    const double pdof = opsi - eyaw + 3.14159265358979323846;
    const double dv = sqrt(ev * ev + ovo * ovo + 2 * ev * ovo * cos(pdof));
    const double dve = ke * dv, dvo = ko * dv;
    if (prot) {
        const double rel = atan2(my - ey, mx - ex);
        const double a_e = rel - eyaw;
        const double a_o = 3.14159265358979323846 + rel - opsi;
        arg_e = pr.lr_speed * dve + angle_coef<double>(pr, a_e);
        arg_o = pr.lr_speed * dvo + angle_coef<double>(pr, a_o);
    } else {
        arg_e = pr.lr1s_speed * dve;
        arg_o = pr.ped_speed * dvo;
    }
}

template <typename R> __device__ __forceinline__ R sigmoid_harm(R cst, R arg) { return (R)1 / ((R)1 + exp_(-cst - arg)); }

// collision probability of one (ego sample i = t+1, obstacle sample t) pair
// (collision_probability.py:205-252, 329-364): 3 means x 3 axis-aligned ego boxes.
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

template <typename R, typename Rect>
__device__ __forceinline__ R nine_rectangles(const Tab<R>& pr, const Rect& rect, R x, R y, R c, R s, R mx, R my, R ex, R ey,
                                             R isx, R isy, R rho) {
    const R hx = pr.hx, hy = pr.hy, rr = pr.rr;
    R prob = 0;
#pragma unroll 1
    for (int km = 0; km < 3; ++km) {
        const R sg = km == 0 ? (R)0 : (km == 1 ? (R)1 : (R)-1);
        const R ux = mx + sg * ex, uy = my + sg * ey;
#pragma unroll 1
        for (int jb = 0; jb < 3; ++jb) {
            const R sb = jb == 0 ? (R)0 : (jb == 1 ? (R)1 : (R)-1);
            const R bx = x + sb * (rr * c), by = y + sb * (rr * s);
            const R xl = ((bx - hx) - ux) * isx, xu = ((bx + hx) - ux) * isx;
            const R yl = ((by - hy) - uy) * isy, yu = ((by + hy) - uy) * isy;
            prob += rect(xl, xu, yl, yu, rho);
        }
    }
    return prob / 3;
}

__device__ __forceinline__ double collision_prob(const Tab<double>& pr, double x, double y, double c, double s, double mx,
                                                 double my, double ex, double ey, double isx, double isy, double rho) {
    return nine_rectangles<double>(pr, RectDouble{}, x, y, c, s, mx, my, ex, ey, isx, isy, rho);
}

__device__ float collision_prob(const Tab<float>& pr, float x, float y, float c, float s, float mx, float my, float ex,
                                float ey, float isx, float isy, float rho) {
    // Node count by |rho|: 0 (rho = 0), 4 (<= 0.35), 6 (<= 0.55), fp64 fallback beyond.  The lanes of a
    // warp hold different (obstacle, timestep) pairs; to keep them on ONE code path every lane uses the
    // largest node count any of them needs (more nodes are never less accurate; rho = 0 gives zero weights).
    const float ar = fabsf(rho);
    const int cls = !(ar <= 0.55f) ? 3 : (ar > 0.35f ? 2 : (rho != 0.0f ? 1 : 0));
    const unsigned act = __activemask();
    const int wcls = __reduce_max_sync(act, cls == 3 ? 0 : cls);
    if (cls == 3) {
        atomicAdd(&g_fallback_evals, 1ull);
        return nine_rectangles<float>(pr, RectFallbackF{}, x, y, c, s, mx, my, ex, ey, isx, isy, rho);
    }
    if (wcls == 0) {
        RectNodesF<0> r;
        return nine_rectangles<float>(pr, r, x, y, c, s, mx, my, ex, ey, isx, isy, rho);
    }
    if (wcls == 1) {
        RectNodesF<4> r;
        r.n = make_rho_nodes<4>(rho);
        return nine_rectangles<float>(pr, r, x, y, c, s, mx, my, ex, ey, isx, isy, rho);
    }
    RectNodesF<6> r;
    r.n = make_rho_nodes<6>(rho);
    return nine_rectangles<float>(pr, r, x, y, c, s, mx, my, ex, ey, isx, isy, rho);
}

// 1e-4 / mahalanobis(ego[i], pos[i-1], inv(cov[i-1]))  (collision_probability.py:281-289)
template <typename R> __device__ __forceinline__ R inv_mahalanobis(R x, R y, R mx, R my, R i00, R i01, R i11) {
    const R dx = x - mx, dy = y - my;
    const R m2 = dx * dx * i00 + dx * dy * i01 + dy * dy * i11;
    return (R)1e-4 / sqrt_(m2);
}

// One deferred (un-gated) evaluation: collision probability of ego sample t+1 against prediction
// sample t, harm of sample t, risk = harm * probability folded into the per-obstacle maxima.
// Kept out of line so that its register demand does not leak into the gated streaming loop.
template <typename R> struct EntryCtx {
    const Tab<R>* tab;
    const Vec4<R>* e0;
    const Vec2<R>* e1;
    Obs<R> ob;
    typename Enc<R>::U* rmax;
    const double* obs_const;
    R* o_prob;
    int On, mahalanobis;
    size_t prob_base;
};

template <typename R> __device__ __noinline__ void run_entry_fn(const EntryCtx<R>& c, unsigned entry) {
    const Tab<R>& tab = *c.tab;
    const Obs<R>& ob = c.ob;
    const int o = entry >> 8, t = entry & 255;
    const int vi = t * ob.O + o;
    const Vec4<R> g0 = ldg4<R>(ob.v0 + vi), g1 = ldg4<R>(ob.v1 + vi);
    const Vec4<R> a1 = c.e0[t + 1], a0 = c.e0[t];
    const Vec2<R> b0 = c.e1[t];
    R prob;
    if (c.mahalanobis) {
        prob = inv_mahalanobis<R>(a1.x, a1.y, g0.x, g0.y, ob.get(OF_I00, o, t), ob.get(OF_I01, o, t), ob.get(OF_I11, o, t));
    } else {
        prob = collision_prob(tab, a1.x, a1.y, a1.z, a1.w, g0.x, g0.y, g0.z, g0.w, ob.get(OF_ISX, o, t), ob.get(OF_ISY, o, t),
                              ob.get(OF_RHO, o, t));
    }
    const double* oc = c.obs_const + o * OC_COUNT;
    const bool prot = oc[OC_PROT] != 0.0;
    R ae, ao;
    harm_args<R>(tab, prot, (R)oc[OC_KE], (R)oc[OC_KO], a0.x, a0.y, a0.z, a0.w, b0.x, b0.y, g0.x, g0.y, g1.w, g1.x, g1.y, g1.z,
                 ae, ao);
    const R he = sigmoid_harm<R>(prot ? tab.lr_const : tab.lr1s_const, ae);
    const R ho = sigmoid_harm<R>(prot ? tab.lr_const : tab.neg_ped_const, ao);
    atomicMax(&c.rmax[o], Enc<R>::enc(he * prob));  // risk_costs.py:88-95
    atomicMax(&c.rmax[c.On + o], Enc<R>::enc(ho * prob));
    if (c.o_prob) c.o_prob[c.prob_base + vi] = prob;
}

struct WarpBest {
    unsigned long long hi, lo;
    double cost;
    int level, scored;
};

template <typename R>
__global__ void __launch_bounds__(kThreads, 2) score_kernel(DevStep st, DevParams pr, DevOut out, BestRec* __restrict__ block_best,
                                                         unsigned int* __restrict__ counters, BestRec* __restrict__ best_out,
                                                         const int* __restrict__ prof_nskip) {
    using U = typename Enc<R>::U;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Ts = st.Tmax;
    const int O = st.has_pred ? st.O : 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    WarpSmem<R> sm(smem_raw + (size_t)warp * warp_smem_bytes<R>(Ts, O), Ts, O);
    const int Cs = st.n_local;  // candidates scored by this call: profiles p_begin + shard_index + k * shard_count
    const int On = O > 0 ? O : 1;
    Obs<R> ob(st.obs, st.O, st.Tp);
    R* o_traj = reinterpret_cast<R*>(out.traj);
    R* o_prob = reinterpret_cast<R*>(out.prob);
    R* o_red = reinterpret_cast<R*>(out.red);
    R* o_cost = reinterpret_cast<R*>(out.cost);

    const bool mahal = pr.mahalanobis != 0;
    __shared__ Tab<R> tab;
    if (tid == 0) fill_tab<R>(tab, pr);
    __syncthreads();

    WarpBest wb{~0ull, ~0ull, 0.0, -1, 0};

    for (;;) {
        int chunk0 = 0;
        if (lane == 0) chunk0 = (int)atomicAdd(&counters[1], (unsigned)kChunk);
        chunk0 = __shfl_sync(kFull, chunk0, 0);
        if (chunk0 >= Cs) break;
        const int chunk1 = min(chunk0 + kChunk, Cs);
        for (int cl = chunk0; cl < chunk1; ++cl) {
            const int kl = cl / st.nd, id = cl - kl * st.nd;
            const int pl = st.shard_index + kl * st.shard_count;  // row of the profile table
            const int p = st.p_begin + pl;
            const int c_dense = p * st.nd + id;
            const double* gp = st.prof + (size_t)pl * PF_COUNT * Ts;
            const double* pm = st.prof_meta + (size_t)pl * PM_COUNT;
            const double ds = pm[PM_DS];
            const bool low = pm[PM_LOW] != 0.0;
            const int T = (int)pm[PM_T];
            const bool vel_ok = pm[PM_VEL_OK] != 0.0;
            const double dT = st.d_list[id];
            const bool skip = ds <= fabs(dT);  // frenet_functions.py:333-334

            // ---- generation + transform: lanes over time ------------------------------------------
            double sv = 0, sd = 0, jl = 0, jq = 0, trav = 0;
            int first_fail = 0x7fffffff;
            {
                const Poly5 q = lateral_poly(st, low, dT, pm[PM_TEND], ds);
                for (int t = lane; t < Ts; t += 32) {
                    const size_t gi = (size_t)cl * Ts + t;
                    if (t < T && !skip) {
                        const TrajPoint tp = traj_point(q, low, gp, Ts, t, st.dt);
                        if (o_traj) {
#pragma unroll
                            for (int f = 0; f < ETP_NUM_FIELDS; ++f) o_traj[(size_t)f * Cs * Ts + gi] = (R)tp.f[f];
                        }
                        sm.xd[t] = tp.f[ETP_F_X];
                        sm.yd[t] = tp.f[ETP_F_Y];
                        sm.e0[t] = Vec4<R>{(R)tp.f[ETP_F_X], (R)tp.f[ETP_F_Y], (R)tp.cos_yaw, (R)tp.sin_yaw};
                        sm.e1[t] = Vec2<R>{(R)tp.f[ETP_F_V], (R)(sizeof(R) == 4 ? tp.wrap : tp.f[ETP_F_YAW])};
                        sv += tp.f[ETP_F_V];
                        sd += fabs(tp.f[ETP_F_D]);
                        jl += tp.f[ETP_F_S_DDD] * tp.f[ETP_F_S_DDD];
                        jq += tp.f[ETP_F_D_DDD] * tp.f[ETP_F_D_DDD];
                        const int cf = curvature_fail(pr, tp.f[ETP_F_V], tp.f[ETP_F_CURV]);
                        if (cf) first_fail = min(first_fail, t * 4 + cf);
                    } else if (o_traj) {
#pragma unroll
                        for (int f = 0; f < ETP_NUM_FIELDS; ++f) o_traj[(size_t)f * Cs * Ts + gi] = (R)0;
                    }
                }
                for (int o = lane; o < 2 * On; o += 32) sm.rmax[o] = Enc<R>::enc((R)-INFINITY);
                __syncwarp();
                if (!skip)
                    for (int t = lane; t + 1 < T; t += 32) {  // calc_trajectory_cost.py:739-757
                        const double dx = sm.xd[t] - sm.xd[t + 1], dy = sm.yd[t] - sm.yd[t + 1];
                        trav += sqrt(dx * dx + dy * dy);
                    }
                sv = warp_sum(sv); sd = warp_sum(sd); jl = warp_sum(jl); jq = warp_sum(jq); trav = warp_sum(trav);
                first_fail = -warp_max(-first_fail);
            }

            // ---- risk rows: lanes over obstacles, loop over time -------------------------------------
            // Lane l owns obstacle ob + (l & (Op-1)) of the current block of <= 32 obstacles; when the
            // block has fewer than 17 obstacles the spare lanes take further timesteps (tt = l / Op).
            // Running maxima over time stay in registers; the obstacle samples of a step are two
            // coalesced 16-byte loads per lane ([t][o] vector planes) and the ego sample is a
            // shared-memory broadcast.
            if (O > 0 && !skip) {
                unsigned* q = sm.queue;
                int q_n = 0;
                R* prob_c = o_prob ? o_prob + (size_t)cl * (Ts - 1) * O : nullptr;
                EntryCtx<R> ectx{&tab, sm.e0, sm.e1, ob, sm.rmax, st.obs_const, o_prob, On, pr.mahalanobis,
                                 (size_t)cl * (Ts - 1) * O};
                for (int ob0 = 0; ob0 < O; ob0 += 32) {
                    const int rem = min(32, O - ob0);
                    int Op = 32;
                    while ((Op >> 1) >= rem) Op >>= 1;  // smallest power of two >= rem
                    const int TT = 32 / Op;
                    const int ol = lane & (Op - 1), tt = lane / Op;
                    const int o = ob0 + ol;
                    const bool o_ok = ol < rem;
                    int n_pred = 0;
                    bool prot = false;
                    R ke = 0, ko = 0;
                    double halflen = 0;
                    if (o_ok) {
                        const double* oc = st.obs_const + o * OC_COUNT;
                        n_pred = st.pred_len[o];
                        prot = oc[OC_PROT] != 0.0;
                        ke = (R)oc[OC_KE];
                        ko = (R)oc[OC_KO];
                        halflen = oc[OC_HALFLEN];
                    }
                    const int t_harm = min(T - 1, n_pred);  // harm samples: t < min(T-1, Tp)   (harm_estimation.py:234)
                    R hm_e = -INFINITY, hm_o = -INFINITY, rm = -INFINITY;
                    for (int t0 = 0; t0 < Ts - 1; t0 += TT) {
                        const int t = t0 + tt;
                        bool defer = false;
                        if (o_ok && t < Ts - 1) {
                            const int vi = t * O + o;
                            if (t < t_harm) {
                                const Vec4<R> g0 = ldg4<R>(ob.v0 + vi), g1 = ldg4<R>(ob.v1 + vi);
                                const Vec4<R> a0 = sm.e0[t];
                                const Vec2<R> b0 = sm.e1[t];
                                // harm sample t: ego[t] against prediction[t] (harm_estimation.py:313-332)
                                R ae, ao;
                                harm_args<R>(tab, prot, ke, ko, a0.x, a0.y, a0.z, a0.w, b0.x, b0.y, g0.x, g0.y, g1.w, g1.x, g1.y,
                                             g1.z, ae, ao);
                                hm_e = rmax_(hm_e, ae);
                                hm_o = rmax_(hm_o, ao);
                                bool gated = true;
                                if (t + 1 < n_pred) {
                                    // probability sample t: ego[t+1] against mean[t], orientation[t+1]
                                    // (collision_probability.py:168-203)
                                    if (mahal) {
                                        gated = false;
                                    } else {
                                        const Vec4<R> a1 = sm.e0[t + 1];
                                        const R x = a1.x, y = a1.y;
                                        R dx = g0.x - x, dy = g0.y - y;
                                        R d2 = dx * dx + dy * dy;
                                        dx = (g0.x + g0.z) - x; dy = (g0.y + g0.w) - y;
                                        d2 = fmin(d2, dx * dx + dy * dy);
                                        dx = (g0.x - g0.z) - x; dy = (g0.y - g0.w) - y;
                                        d2 = fmin(d2, dx * dx + dy * dy);
                                        if (sizeof(R) == 8) {
                                            gated = sqrt_(d2) > (R)5.0;
                                        } else {
                                            gated = d2 > (R)25.0;
                                            if (abs_(d2 - (R)25) < (R)2e-3 + (R)1e-5 * (abs_(x) + abs_(y))) {
                                                // fp32 cannot decide the 5 m gate here: repeat it in fp64 from the exact inputs
                                                const double xe = sm.xd[t + 1], ye = sm.yd[t + 1];
                                                const double* rp = st.raw_pos + ((size_t)o * st.Tp + t) * 2;
                                                const double y1 = st.raw_yaw[(size_t)o * st.Tp + t + 1];
                                                const double len = 2 * halflen;
                                                const double fx = cos(y1) * len / 2, fy = sin(y1) * len / 2;
                                                double ax = rp[0] - xe, ay = rp[1] - ye;
                                                double m = sqrt(ax * ax + ay * ay);
                                                ax = (rp[0] + fx) - xe; ay = (rp[1] + fy) - ye;
                                                m = fmin(m, sqrt(ax * ax + ay * ay));
                                                ax = (rp[0] - fx) - xe; ay = (rp[1] - fy) - ye;
                                                m = fmin(m, sqrt(ax * ax + ay * ay));
                                                gated = m > 5.0;
                                            }
                                        }
                                    }
                                }
                                if (gated) rm = rmax_(rm, (R)0);  // harm * 0.0
                                else defer = true;
                            }
                            if (!defer && prob_c) prob_c[vi] = (R)0;
                        }
                        const unsigned m = __ballot_sync(kFull, defer);
                        if (m) {
                            if (defer) q[q_n + __popc(m & ((1u << lane) - 1))] = ((unsigned)o << 8) | (unsigned)t;
                            q_n += __popc(m);
                            __syncwarp();
                            if (q_n >= 32) {
                                q_n -= 32;
                                run_entry_fn<R>(ectx, q[q_n + lane]);
                                __syncwarp();
                            }
                        }
                    }
                    // lanes that share an obstacle (different tt) combine their maxima
                    for (int off = Op; off < 32; off <<= 1) {
                        hm_e = rmax_(hm_e, __shfl_xor_sync(kFull, hm_e, off));
                        hm_o = rmax_(hm_o, __shfl_xor_sync(kFull, hm_o, off));
                        rm = rmax_(rm, __shfl_xor_sync(kFull, rm, off));
                    }
                    if (o_ok && tt == 0) {
                        sm.hmax[o] = hm_e;
                        sm.hmax[On + o] = hm_o;
                        if (rm > (R)-INFINITY) {
                            atomicMax(&sm.rmax[o], Enc<R>::enc(rm));
                            atomicMax(&sm.rmax[On + o], Enc<R>::enc(rm));
                        }
                    }
                }
                if (lane < q_n) run_entry_fn<R>(ectx, q[lane]);  // drain: entries live at [0, q_n)
                __syncwarp();
            } else if (O > 0 && o_prob) {
                for (int i = lane; i < O * (Ts - 1); i += 32) o_prob[(size_t)cl * O * (Ts - 1) + i] = (R)0;
            }

            // ---- principles (lanes over obstacles), cost, level, key --------------------------------
            double c[ETP_NUM_COST];
#pragma unroll
            for (int k = 0; k < ETP_NUM_COST; ++k) c[k] = 0.0;
            int level = ETP_LEVEL_SKIPPED, reason = ETP_REASON_NONE;
            double sum_e = 0, sum_o = 0, sum_abs = 0, resp = 0, mm = -INFINITY, max_o = -INFINITY;
            for (int o = lane; o < O; o += 32) {
                R re = 0, ro = 0, he = 0, ho = 0;
                if (!skip) {
                    const double* oc = st.obs_const + o * OC_COUNT;
                    const bool prot = oc[OC_PROT] != 0.0;
                    re = Enc<R>::dec(sm.rmax[o]);
                    ro = Enc<R>::dec(sm.rmax[On + o]);
                    he = sigmoid_harm<R>(prot ? tab.lr_const : tab.lr1s_const, sm.hmax[o]);
                    ho = sigmoid_harm<R>(prot ? tab.lr_const : tab.neg_ped_const, sm.hmax[On + o]);
                    sum_e += (double)re;
                    sum_o += (double)ro;
                    sum_abs += fabs((double)re - (double)ro);
                    resp -= oc[OC_RESP] * (double)ro;                                // risk_costs.py:250-253
                    mm = fmax(mm, (double)he * ((double)re < 10e-10 ? 1.0 : 0.0));  // risk_costs.py:205-206 (Q3)
                    mm = fmax(mm, (double)ho * ((double)ro < 10e-10 ? 1.0 : 0.0));
                    max_o = fmax(max_o, (double)ro);
                }
                if (o_red) {
                    const size_t g = (size_t)cl * O + o, plane = (size_t)Cs * O;
                    o_red[ETP_R_EGO_RISK * plane + g] = re;
                    o_red[ETP_R_OBST_RISK * plane + g] = ro;
                    o_red[ETP_R_EGO_HARM * plane + g] = he;
                    o_red[ETP_R_OBST_HARM * plane + g] = ho;
                }
            }
            if (!skip) {
                sum_e = warp_sum(sum_e); sum_o = warp_sum(sum_o); sum_abs = warp_sum(sum_abs); resp = warp_sum(resp);
                mm = warp_max(mm); max_o = warp_max(max_o);
                const double bd = st.bd_harm ? st.bd_harm[c_dense] : 0.0;
                const int ext = st.ext_level ? (int)st.ext_level[c_dense] : 10;
                if (O > 0) {
                    c[ETP_C_BAYES] = (sum_e + sum_o + bd) / (double)(O * 2);  // risk_costs.py:148-150
                    c[ETP_C_EQUALITY] = sum_abs / (double)O;                  // :176-178
                    c[ETP_C_MAXIMIN] = pow(fmax(mm, bd), 10.0);               // :208
                    c[ETP_C_EGO] = sum_e + bd;                                // :226
                    c[ETP_C_RESPONSIBILITY] = resp;
                }
                const bool risk_in_cost = pr.w[ETP_W_RISK_COST] > 0.0 && st.has_pred && pr.planner_mode == ETP_MODE_RISK;
                if (risk_in_cost)  // calc_trajectory_cost.py:130-136
                    c[ETP_C_RISK_TOTAL] = pr.w[ETP_W_BAYES] * c[ETP_C_BAYES] + pr.w[ETP_W_EQUALITY] * c[ETP_C_EQUALITY] +
                                          pr.w[ETP_W_MAXIMIN] * c[ETP_C_MAXIMIN] +
                                          pr.w[ETP_W_RESPONSIBILITY] * pr.w[ETP_W_BAYES] * c[ETP_C_RESPONSIBILITY] +
                                          pr.w[ETP_W_EGO] * c[ETP_C_EGO];
                const double mean_v = sv / (double)T;
                switch (st.vel_mode) {  // calc_trajectory_cost.py:438-519 outcome forms
                    case 0: c[ETP_C_VELOCITY] = fabs(st.vel_target - mean_v); break;
                    case 1: c[ETP_C_VELOCITY] = 30.0 - mean_v; break;
                    case 2: c[ETP_C_VELOCITY] = mean_v; break;
                    default: c[ETP_C_VELOCITY] = 0.0;
                }
                c[ETP_C_DIST_TO_GLOBAL_PATH] = sd / (double)T;  // :726-736
                c[ETP_C_LON_JERK] = jl / (double)T;             // :418-435
                c[ETP_C_LAT_JERK] = jq / (double)T;
                c[ETP_C_TRAVELLED_DIST] = trav;
                c[ETP_C_FACTOR] = st.factor;
                // validity level (validity_checks.py:31-122)
                const int cf = first_fail == 0x7fffffff ? 0 : (first_fail & 3);
                if (!vel_ok) { level = 0; reason = ETP_REASON_VELOCITY; }
                else if (cf) { level = 0; reason = cf == 1 ? ETP_REASON_CURV_LVC : ETP_REASON_CURV_HVC; }
                else if (ext == 1) { level = 1; reason = ETP_REASON_COLLISION; }
                else if (st.has_pred ? (bd != 0.0) : (ext == 2)) { level = 2; reason = ETP_REASON_BOUNDARIES; }
                else if (pr.planner_mode == ETP_MODE_RISK && st.has_pred && O > 0 &&
                         (sum_e > pr.max_risk || max_o > pr.max_risk)) { level = 3; reason = ETP_REASON_MAX_RISK; }
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
                c[ETP_C_TOTAL] = st.factor * cost;
                const unsigned long long hi = make_key_hi(level, c[ETP_C_TOTAL]);
                const unsigned long long lo = (unsigned long long)c_dense;
                if (hi < wb.hi || (hi == wb.hi && lo < wb.lo)) {
                    wb.hi = hi; wb.lo = lo; wb.cost = c[ETP_C_TOTAL]; wb.level = level;
                }
                wb.scored += 1;
            }
            if (o_cost && lane < ETP_NUM_COST) {
                double v = 0.0;
#pragma unroll
                for (int k = 0; k < ETP_NUM_COST; ++k) v = (lane == k) ? c[k] : v;
                o_cost[(size_t)lane * Cs + cl] = (R)v;
            }
            if (lane == 0) {
                if (out.level) out.level[cl] = (uint8_t)level;
                if (out.reason) out.reason[cl] = (uint8_t)reason;
            }
            __syncwarp();
        }
    }

    // ---- block arg-min, then grid arg-min by the last block to finish ---------------------------
    {
        __shared__ BestRec wbest[kWarps];
        __shared__ bool is_last;
        if (lane == 0) {
            wbest[warp].key_hi = wb.hi; wbest[warp].key_lo = wb.lo; wbest[warp].cost = wb.cost;
            wbest[warp].level = wb.level; wbest[warp].n_scored = wb.scored;
        }
        __syncthreads();
        if (tid == 0) {
            BestRec b = wbest[0];
            for (int w = 1; w < kWarps; ++w) {
                const int n = b.n_scored + wbest[w].n_scored;
                if (wbest[w].key_hi < b.key_hi || (wbest[w].key_hi == b.key_hi && wbest[w].key_lo < b.key_lo)) b = wbest[w];
                b.n_scored = n;
            }
            b.index = b.key_hi == ~0ull ? -1 : (int)b.key_lo;
            b.list_index = -1;
            block_best[blockIdx.x] = b;
            __threadfence();
            const unsigned done = atomicAdd(&counters[0], 1u);
            is_last = (done == gridDim.x - 1);
        }
        __syncthreads();
        if (is_last) {
            __threadfence();
            unsigned long long hi = ~0ull, lo = ~0ull;
            double cst = 0.0;
            int lvl = -1, scored = 0;
            for (int b = tid; b < (int)gridDim.x; b += kThreads) {
                const volatile BestRec* r = block_best + b;
                const unsigned long long bhi = r->key_hi, blo = r->key_lo;
                scored += r->n_scored;
                if (bhi < hi || (bhi == hi && blo < lo)) { hi = bhi; lo = blo; cst = r->cost; lvl = r->level; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const unsigned long long ohi = __shfl_xor_sync(kFull, hi, o), olo = __shfl_xor_sync(kFull, lo, o);
                const double oc = __shfl_xor_sync(kFull, cst, o);
                const int ol = __shfl_xor_sync(kFull, lvl, o);
                scored += __shfl_xor_sync(kFull, scored, o);
                if (ohi < hi || (ohi == hi && olo < lo)) { hi = ohi; lo = olo; cst = oc; lvl = ol; }
            }
            if (lane == 0) {
                wbest[warp].key_hi = hi; wbest[warp].key_lo = lo; wbest[warp].cost = cst; wbest[warp].level = lvl;
                wbest[warp].n_scored = scored;
            }
            __syncthreads();
            if (tid == 0) {
                BestRec b = wbest[0];
                for (int w = 1; w < kWarps; ++w) {
                    const int n = b.n_scored + wbest[w].n_scored;
                    if (wbest[w].key_hi < b.key_hi || (wbest[w].key_hi == b.key_hi && wbest[w].key_lo < b.key_lo)) b = wbest[w];
                    b.n_scored = n;
                }
                b.index = b.key_hi == ~0ull ? -1 : (int)b.key_lo;
                b.list_index = -1;
                if (b.index >= 0) {
                    // rank among non-skipped candidates of this range = position in the reference's fp_list
                    const int pb = b.index / st.nd, idb = b.index % st.nd;
                    int li = 0;
                    for (int pp = st.p_begin; pp < pb; ++pp) li += st.nd - prof_nskip[pp - st.p_begin];
                    const double dsb = st.prof_meta[(size_t)(pb - st.p_begin) * PM_COUNT + PM_DS];
                    for (int id = 0; id < idb; ++id) li += (dsb <= fabs(st.d_list[id])) ? 0 : 1;
                    b.list_index = li;
                }
                *best_out = b;
                counters[0] = 0;  // re-arm both counters for the next launch
                counters[1] = 0;
            }
        }
    }
}

// number of skipped lateral targets per profile (frenet_functions.py:333-334)
__global__ void profile_skip_kernel(DevStep st, int* __restrict__ nskip) {
    const int pl = blockIdx.x * blockDim.x + threadIdx.x;
    if (pl >= st.p_end - st.p_begin) return;
    const double ds = st.prof_meta[(size_t)pl * PM_COUNT + PM_DS];
    int n = 0;
    for (int id = 0; id < st.nd; ++id) n += (ds <= fabs(st.d_list[id])) ? 1 : 0;
    nskip[pl] = n;
}

// one candidate, all 13 fields in fp64 -> [ETP_NUM_FIELDS][Tmax]; requires its profile in st.prof
__global__ void trajectory_kernel(DevStep st, int c_dense, double* __restrict__ fields, int* __restrict__ n_out) {
    const int p = c_dense / st.nd, id = c_dense % st.nd;
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

// ---------------------------------------------------------------------------------------------
// host-side launchers (called from etp_abi.cpp)
// ---------------------------------------------------------------------------------------------
cudaError_t launch_pack_obstacles(int precision, int O, int Tp, const int* pred_len, const double* pos, const double* cov,
                                  const double* yaw, const double* vel, const double* length, const double* mass,
                                  const int* prot, const int* resp, double veh_m, void* tile, double* oconst,
                                  cudaStream_t stream) {
    const int n = O * Tp;
    if (n <= 0) return cudaSuccess;
    const int blocks = (n + 127) / 128;
    if (precision == ETP_PRECISION_FP64)
        pack_obstacles_kernel<double><<<blocks, 128, 0, stream>>>(O, Tp, pred_len, pos, cov, yaw, vel, length, mass, prot, resp,
                                                                  veh_m, (double*)tile, oconst);
    else
        pack_obstacles_kernel<float><<<blocks, 128, 0, stream>>>(O, Tp, pred_len, pos, cov, yaw, vel, length, mass, prot, resp,
                                                                 veh_m, (float*)tile, oconst);
    return cudaGetLastError();
}

cudaError_t launch_profiles(const DevStep& st, const DevParams& pr, int* nskip, cudaStream_t stream) {
    const int n_prof = st.p_end - st.p_begin;
    if (n_prof <= 0) return cudaSuccess;
    const size_t smem = sizeof(double) * 7 * st.Tmax;
    profile_kernel<<<n_prof, 128, smem, stream>>>(st, pr);
    profile_skip_kernel<<<(n_prof + 127) / 128, 128, 0, stream>>>(st, nskip);
    return cudaGetLastError();
}

template <typename R>
static cudaError_t launch_score_t(const DevStep& st, const DevParams& pr, const DevOut& out, BestRec* block_best,
                                  unsigned* counters, BestRec* best_out, const int* nskip, int num_sms, int* grid_out,
                                  cudaStream_t stream) {
    const int O = st.has_pred ? st.O : 0;
    const size_t smem = kWarps * warp_smem_bytes<R>(st.Tmax, O);
    cudaError_t e = cudaFuncSetAttribute(score_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, score_kernel<R>, kThreads, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    // persistent grid: every SM fully occupied, never more CTAs than there are work chunks
    const long long chunks = ((long long)st.n_local + kChunk - 1) / kChunk;
    long long grid = (long long)num_sms * per_sm;
    const long long need = (chunks + kWarps - 1) / kWarps;
    if (grid > need) grid = need;
    if (grid > kMaxGrid) grid = kMaxGrid;
    if (grid < 1) grid = 1;
    *grid_out = (int)grid;
    score_kernel<R><<<(int)grid, kThreads, smem, stream>>>(st, pr, out, block_best, counters, best_out, nskip);
    return cudaGetLastError();
}

cudaError_t launch_score(int precision, const DevStep& st, const DevParams& pr, const DevOut& out, BestRec* block_best,
                         unsigned* counters, BestRec* best_out, const int* nskip, int num_sms, int* grid_out,
                         cudaStream_t stream) {
    if (precision == ETP_PRECISION_FP64)
        return launch_score_t<double>(st, pr, out, block_best, counters, best_out, nskip, num_sms, grid_out, stream);
    return launch_score_t<float>(st, pr, out, block_best, counters, best_out, nskip, num_sms, grid_out, stream);
}

cudaError_t read_debug_counters(unsigned long long* out, int n) {
    unsigned long long v = 0;
    cudaError_t e = cudaMemcpyFromSymbol(&v, g_fallback_evals, sizeof(v));
    if (n > 0) out[0] = v;
    return e;
}

cudaError_t launch_trajectory(const DevStep& st, int c_dense, double* fields, int* n_out, cudaStream_t stream) {
    trajectory_kernel<<<1, 64, 0, stream>>>(st, c_dense, fields, n_out);
    return cudaGetLastError();
}

}  // namespace etp
