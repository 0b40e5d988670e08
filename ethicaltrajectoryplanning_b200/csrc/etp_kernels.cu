// Input-side kernels of the B200-native candidate-trajectory path (sm_100a).
//
//   pack_obstacles_kernel : predictions (SoA fp64) -> packed obstacle tile of R (layout in etp_device.cuh)
//   profile_kernel        : per (v,t) profile stage in fp64: quartic, spline lookup, chained
//                           non-uniform gradients, reference yaw / curvature / curvature'
//                           (frenet_functions.py:258-325)
//   trajectory_kernel     : one candidate recomputed in fp64 for the host (frenet_planner.py:564-576)
//
// The fused per-candidate scoring kernel lives in etp_score.cu.
#include "etp_device.cuh"
#include "etp_launch.h"

namespace etp {

constexpr unsigned kFull = 0xffffffffu;

// ---------------------------------------------------------------------------------------------
// pack_obstacles_kernel: one thread per (obstacle, prediction sample)
// ---------------------------------------------------------------------------------------------
template <typename R>
__device__ __forceinline__ void pack_sample(const PackArgs& a, int i) {
    const int O = a.O, Tp = a.Tp;
    const int* __restrict__ pred_len = a.pred_len;
    const double* __restrict__ pos = a.pos;
    const double* __restrict__ cov = a.cov;
    const double* __restrict__ yaw = a.yaw;
    const double* __restrict__ vel = a.vel;
    const double* __restrict__ length = a.length;
    const double* __restrict__ width = a.width;
    const double* __restrict__ mass = a.mass;
    const int* __restrict__ prot = a.prot;
    const int* __restrict__ resp = a.resp;
    const double veh_m = a.veh_m, ox = a.ox, oy = a.oy;
    void* tile_base = a.tile;
    double* __restrict__ oconst = a.oconst;
    const int o = i / Tp, t = i % Tp;
    const int n = pred_len[o];
    const Tile<R> tile(tile_base, O, Tp);
    Vec4<R>* G = const_cast<Vec4<R>*>(tile.G) + 2 * ((size_t)t * O + o);
    Vec4<R>* P = const_cast<Vec4<R>*>(tile.P) + 2 * ((size_t)t * O + o);
    float4* N = const_cast<float4*>(tile.N) + kNodeVecs * ((size_t)t * O + o);
    const double PI = 3.14159265358979323846;
    double mx = 0, my = 0, ex = 0, ey = 0, vo = 0, cpsi = 0, spsi = 0, psi = 0, wrap = 0;
    double isx = 0, isy = 0, rho = 0, i00 = 0, i01 = 0, i11 = 0;
    int cls = RC_ZERO;
    float nodes[4 * kNodeVecs];
    for (int k = 0; k < 4 * kNodeVecs; ++k) nodes[k] = 0.0f;
    if (t < n) {
        const double* p = pos + ((size_t)o * Tp + t) * 2;
        const double* c = cov + ((size_t)o * Tp + t) * 4;
        // tile coordinates are relative to the step origin: fp32 keeps millimetres next to the ego instead of 0.06 mm
        // ulps at 600 m; only differences of positions are ever used downstream
        mx = p[0] - ox;
        my = p[1] - oy;
        if (t + 1 < n) {
            const double y1 = yaw[(size_t)o * Tp + t + 1];
            // np.stack((cos, sin)) * length / 2   (collision_probability.py:175)
            ex = cos(y1) * length[o] / 2;
            ey = sin(y1) * length[o] / 2;
        }
        double c00 = c[0], c01 = c[1], c10 = c[2], c11 = c[3];
        if (c00 == 0.0 && c01 == 0.0 && c10 == 0.0 && c11 == 0.0) {  // collision_probability.py:210-212
            c00 = 0.1; c11 = 0.1;
        }
        const double sx = sqrt(c00), sy = sqrt(c11);
        isx = 1.0 / sx;
        isy = 1.0 / sy;
        rho = c10 / sy / sx;  // mvnun takes the correlation from the lower triangle
        psi = yaw[(size_t)o * Tp + t];
        vo = vel[(size_t)o * Tp + t];
        cpsi = cos(psi);
        spsi = sin(psi);
        wrap = ceil((psi - PI) / (2 * PI));
        // Mahalanobis mode inverts cov_list itself (collision_probability.py:279), not the 0.1*I stand-in
        const double r00 = c[0], r01 = c[1], r10 = c[2], r11 = c[3];
        const double det = r00 * r11 - r01 * r10;
        i00 = r11 / det;
        i01 = -(r01 + r10) / det;
        i11 = r00 / det;
        // low parts of the mean: exact mx = (float)mx + nodes[18] (fp32 tile; see nine_rectangles in etp_score.cu)
        nodes[18] = (float)(mx - (double)(float)mx);
        nodes[19] = (float)(my - (double)(float)my);
        // nodes of the rho integral (fp32 path, etp_bvn.cuh): theta_i on [0, asin rho]
        const float rf = (float)rho;
        const float ar = fabsf(rf);
        cls = !(ar <= 0.55f) ? RC_FP64 : (ar > 0.35f ? RC_N6 : (rf != 0.0f ? RC_N4 : RC_ZERO));
        if (cls == RC_N4 || cls == RC_N6) {
            const int nn = cls == RC_N4 ? 4 : 6;
            const double gx4[2] = {0.3399810435848563, 0.8611363115940526};
            const double gw4[2] = {0.6521451548625461, 0.3478548451374538};
            const double gx6[3] = {0.2386191860831969, 0.6612093864662645, 0.9324695142031521};
            const double gw6[3] = {0.4679139345726910, 0.3607615730481386, 0.1713244923791704};
            const double asr = asin((double)rf);
            for (int k = 0; k < nn; ++k) {
                const double g = nn == 4 ? gx4[k >> 1] : gx6[k >> 1];
                const double w = nn == 4 ? gw4[k >> 1] : gw6[k >> 1];
                const double th = asr * (0.5 + ((k & 1) ? 0.5 : -0.5) * g);
                const double sn = sin(th);
                nodes[k] = (float)sn;
                nodes[6 + k] = (float)(1.4426950408889634 / (1.0 - sn * sn));
                nodes[12 + k] = (float)(w * asr * 0.07957747154594767);  // 1 / (4 pi)
            }
        }
    }
    G[0] = Vec4<R>{(R)mx, (R)my, (R)ex, (R)ey};
    // fp32 tile: which (half plane of the ego seen from the obstacle, sign of dy) combinations give an un-wrapped
    // impact angle pi + rel - psi inside [-pi, pi] (quirk Q4; derivation at Logits::coefs in etp_score.cu).
    // bit (pos << 1 | a): pos = ego on the obstacle's clockwise side (angle >= 0), a = (dy >= 0)
    int code = 0;
    {
        const bool u = spsi >= 0.0;
        for (int pos = 0; pos < 2; ++pos)
            for (int a = 0; a < 2; ++a) {
                bool valid;
                if (pos) valid = (wrap == 0.0) ? !(a && !u) : (wrap == 1.0 ? (a && !u) : false);
                else valid = (wrap == 0.0) ? (!a && u) : (wrap == 1.0 ? !(!a && u) : false);
                code |= valid ? (1 << (pos * 2 + a)) : 0;
            }
    }
    G[1] = Vec4<R>{(R)vo, (R)cpsi, (R)spsi, sizeof(R) == 4 ? (R)__int_as_float(code) : (R)psi};
#ifdef ETP_TMA_STAGE
    {
        Vec4<R>* Gom = const_cast<Vec4<R>*>(tile.Gom) + 2 * ((size_t)o * Tp + t);
        Gom[0] = G[0];
        Gom[1] = G[1];
    }
#endif
    P[0] = Vec4<R>{(R)isx, (R)isy, (R)rho, (R)cls};
    P[1] = Vec4<R>{(R)i00, (R)i01, (R)i11, (R)0};
    for (int k = 0; k < kNodeVecs; ++k) N[k] = make_float4(nodes[4 * k], nodes[4 * k + 1], nodes[4 * k + 2], nodes[4 * k + 3]);
    if (t == 0) {
        const double mo = mass[o];
        const double ke = mo / (veh_m + mo), ko = veh_m / (veh_m + mo);  // harm_estimation.py:327-328
        oconst[o * OC_COUNT + OC_KE] = ke;
        oconst[o * OC_COUNT + OC_KO] = ko;
        oconst[o * OC_COUNT + OC_PROT] = prot[o] ? 1.0 : 0.0;
        oconst[o * OC_COUNT + OC_RESP] = (double)resp[o];
        oconst[o * OC_COUNT + OC_HALFLEN] = length[o] / 2;
        oconst[o * OC_COUNT + OC_HALFWID] = width[o] / 2;
        double step = 0.0;  // reach of the swept collision test (etp_score.cu)
        for (int k = 0; k + 1 < n; ++k) {
            const double* q = pos + ((size_t)o * Tp + k) * 2;
            step = fmax(step, fabs(q[2] - q[0]) + fabs(q[3] - q[1]));
        }
        oconst[o * OC_COUNT + OC_MAXSTEP] = step;
        const_cast<Vec4<R>*>(tile.C)[o] = Vec4<R>{(R)ke, (R)ko, (R)(prot[o] ? 1 : 0), (R)resp[o]};
    }
}

template <typename R> __global__ void pack_obstacles_kernel(const PackArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.O * a.Tp) pack_sample<R>(a, i);
}

// scenario sweep: blockIdx.y = scenario
template <typename R> __global__ void pack_obstacles_batch_kernel(const PackArgs* __restrict__ args) {
    const PackArgs a = args[blockIdx.y];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < a.O * a.Tp) pack_sample<R>(a, i);
}

__device__ __forceinline__ double np_gradient(const double* f, const double* s, int i, int n);

// ---------------------------------------------------------------------------------------------
// enrich_kernel: one block per obstacle (prediction_helpers.py:75-135, responsibility.py:57-89), fp64
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void enrich_body(int O, int Tp, const int* __restrict__ pred_len, const double* __restrict__ pos,
                                            const double* __restrict__ raw_len, const double* __restrict__ raw_wid,
                                            const double* __restrict__ ori0, const double* __restrict__ vel0, double dt,
                                            double margin_l, double margin_w, double ego_x, double ego_y, double ego_ori, bool wrap_view_cone,
                                            double* __restrict__ yaw, double* __restrict__ vel, double* __restrict__ length,
                                            double* __restrict__ width, int* __restrict__ resp, double* sm, const int o) {
    const int tid = threadIdx.x;
    const int n = pred_len[o];
    double* s_x = sm;
    double* s_y = sm + Tp;
    double* s_t = sm + 2 * Tp;
    const double* p = pos + (size_t)o * Tp * 2;
    for (int t = tid; t < n; t += blockDim.x) {
        s_x[t] = p[2 * t];
        s_y[t] = p[2 * t + 1];
        s_t[t] = 0.0 + t * dt;  // t = [0.0 + i * scenario.dt ...]  (:108)
    }
    __syncthreads();
    double* yo = yaw + (size_t)o * Tp;
    double* vo = vel + (size_t)o * Tp;
    if (n == 1) {  // :103-105
        if (tid == 0) {
            yo[0] = ori0[o];
            vo[0] = vel0[o];
        }
    } else {
        // dx = np.gradient(x, t), dy = np.gradient(y, t)  (:113-114); all(dx < 0.0001) and all(dy < 0.0001)  (:117, no abs)
        int still = 1;
        for (int t = tid; t < n; t += blockDim.x) {
            const double dx = np_gradient(s_x, s_t, t, n), dy = np_gradient(s_y, s_t, t, n);
            if (!(dx < 0.0001) || !(dy < 0.0001)) still = 0;
            vo[t] = sqrt(dx * dx + dy * dy);  // :126
            yo[t] = atan2(dy, dx);            // :123 (overwritten below if the obstacle barely moves)
        }
        still = __syncthreads_and(still);
        if (still)
            for (int t = tid; t < n; t += blockDim.x) yo[t] = ori0[o];  // :118-119
    }
    for (int t = n + tid; t < Tp; t += blockDim.x) {
        yo[t] = 0.0;
        vo[t] = 0.0;
    }
    if (tid == 0) {
        length[o] = raw_len[o] + margin_l;  // :131-134
        width[o] = raw_wid[o] + margin_w;
        // check_if_inside180view (responsibility.py:78-89)
        const double ang = atan2(p[1] - ego_y, p[0] - ego_x);
        const double quarter = 3.14159265358979323846 / 4;
        double rel = ang - ego_ori;  // the reference compares ang with ego_ori -+ pi/4 as they are (quirk Q9)
        if (wrap_view_cone) rel = remainder(rel, 2.0 * 3.14159265358979323846);  // ETP_RELAX_Q9_VIEW_CONE: modulo 2 pi
        resp[o] = wrap_view_cone ? ((-quarter <= rel && rel <= quarter) ? 0 : 1)
                                 : ((ego_ori - quarter <= ang && ang <= ego_ori + quarter) ? 0 : 1);
    }
}

__global__ void __launch_bounds__(128) enrich_kernel(int O, int Tp, const int* __restrict__ pred_len, const double* __restrict__ pos,
                                                     const double* __restrict__ raw_len, const double* __restrict__ raw_wid,
                                                     const double* __restrict__ ori0, const double* __restrict__ vel0, double dt,
                                                     double margin_l, double margin_w, double ego_x, double ego_y, double ego_ori, bool wrap_view_cone,
                                                     double* __restrict__ yaw, double* __restrict__ vel, double* __restrict__ length,
                                                     double* __restrict__ width, int* __restrict__ resp) {
    extern __shared__ double sm[];
    enrich_body(O, Tp, pred_len, pos, raw_len, raw_wid, ori0, vel0, dt, margin_l, margin_w, ego_x, ego_y, ego_ori, wrap_view_cone, yaw, vel,
                length, width, resp, sm, blockIdx.x);
}

// the same with its arguments in device memory (planning cycle as a captured graph: the launch never changes, the record does)
__global__ void __launch_bounds__(128) enrich_args_kernel(const EnrichArgs* __restrict__ args) {
    extern __shared__ double sm[];
    const EnrichArgs a = *args;
    if ((int)blockIdx.x >= a.O) return;
    enrich_body(a.O, a.Tp, a.pred_len, a.pos, a.raw_len, a.raw_wid, a.ori0, a.vel0, a.dt, a.margin_l, a.margin_w, a.ego_x, a.ego_y, a.ego_ori,
                a.wrap_view_cone != 0, a.yaw, a.vel, a.length, a.width, a.resp, sm, blockIdx.x);
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

__device__ __forceinline__ void profile_body(const DevStep& st, const DevParams& pr, int* __restrict__ nskip, double* sm, int block) {
    const int Tmax = st.Tmax;
    double* s_s = sm;
    double* s_x = sm + Tmax;
    double* s_y = sm + 2 * Tmax;
    double* s_a = sm + 3 * Tmax;  // dx   -> dddx
    double* s_b = sm + 4 * Tmax;  // dy   -> dddy
    double* s_c = sm + 5 * Tmax;  // ddx
    double* s_d = sm + 6 * Tmax;  // ddy
    const int p = st.p_begin + block;
    const int iv = p / st.nt, it = p % st.nt;
    const double vT = st.v_list[iv], tT = st.t_list[it];
    const int T = st.nsamp[it];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* pf = st.prof + (size_t)block * PF_COUNT * Tmax;

    // quartic_polynomial(xs=c_s, vxs=c_s_d, axs=c_s_dd, vxe=vT, axe=0, T=tT-dt)  (frenet_functions.py:258-260)
    const double Te = tT - st.dt;
    const double a0 = st.c_s, a1 = st.c_s_d, a2 = st.c_s_dd / 2.0;
    const double b1 = vT - a1 - 2 * a2 * Te, b2 = 0.0 - 2 * a2;
    const double a3 = (3.0 * b1 - b2 * Te) / (3.0 * Te * Te);
    const double a4 = (b2 * Te - 2.0 * b1) / (4.0 * Te * Te * Te);

    int vel_ok = 1, acc_ok = 1;
    for (int t = tid; t < T; t += blockDim.x) {
        const double tt = t * st.dt;
        const double t2 = tt * tt, t3 = t2 * tt, t4 = t2 * t2;
        const double s = a0 + a1 * tt + a2 * t2 + a3 * t3 + a4 * t4;
        const double sd = a1 + 2 * a2 * tt + 3 * a3 * t2 + 4 * a4 * t3;
        s_s[t] = s;
        pf[PF_S * Tmax + t] = s;
        pf[PF_S_D * Tmax + t] = sd;
        const double sdd = 2 * a2 + 6 * a3 * tt + 12 * a4 * t2;
        pf[PF_S_DD * Tmax + t] = sdd;
        pf[PF_S_DDD * Tmax + t] = 6 * a3 + 24 * a4 * tt;
        if (!(fabs(sd) <= pr.v_max)) vel_ok = 0;  // validity_checks.py:158
        if ((pr.relaxed & ETP_RELAX_Q2_ACCELERATION) && !(fabs(sdd) <= pr.a_max)) acc_ok = 0;
    }
    vel_ok = __syncthreads_and(vel_ok);
    acc_ok = __syncthreads_and(acc_ok);
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
        // :316-325
        const double z = dx * ddy - ddx * dy;
        const double z_d = dx * dddy - dddx * dy;
        const double q = dx * dx + dy * dy;
        const double n = pow(q, 1.5);
        const double curv = (pr.relaxed & ETP_RELAX_Q1_CURVATURE) ? z / n : (dx * ddy - ddx * dy) / (dx * dx + pow(dy * dy, 1.5));
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
        double* pm = st.prof_meta + (size_t)block * PM_COUNT;
        pm[PM_DS] = s_s[T - 1] - s_s[0];                                                  // :327-328
        pm[PM_LOW] = (fabs(st.c_s_d) < st.v_thr || fabs(vT) < st.v_thr) ? 1.0 : 0.0;  // :248-251
        pm[PM_T] = (double)T;
        pm[PM_VEL_OK] = vel_ok ? (acc_ok ? 1.0 : -1.0) : 0.0;  // 1 fine, 0 velocity, -1 acceleration (relaxed Q2 only)
        pm[PM_TEND] = Te;
    }
    // number of skipped lateral targets of this profile (frenet_functions.py:333-334): rank of the winner in fp_list
    {
        const double ds = s_s[T - 1] - s_s[0];
        int n = 0;
        for (int id = tid; id < st.nd; id += blockDim.x) n += (ds <= fabs(st.d_list[id])) ? 1 : 0;
        __shared__ int s_n;
        if (tid == 0) s_n = 0;
        __syncthreads();
        if (n) atomicAdd(&s_n, n);
        __syncthreads();
        if (tid == 0) nskip[block] = s_n;
    }
}

__global__ void __launch_bounds__(128) profile_kernel(const __grid_constant__ DevStep st, const __grid_constant__ DevParams pr, int* __restrict__ nskip) {
    extern __shared__ double sm[];
    profile_body(st, pr, nskip, sm, blockIdx.x);
}

// scenario sweep: blockIdx.y = scenario, blockIdx.x = profile of that scenario
__global__ void __launch_bounds__(128) profile_batch_kernel(const DevStep* __restrict__ steps, const __grid_constant__ DevParams pr) {
    extern __shared__ double sm[];
    const DevStep& st = steps[blockIdx.y];
    if ((int)blockIdx.x >= st.p_end - st.p_begin) return;
    profile_body(st, pr, st.nskip, sm, blockIdx.x);
}

// Planning cycle (etp_plan_cycle): everything that precedes the fused kernel in ONE launch.  Blocks [0, n_prof) generate the
// longitudinal profiles; block n_prof + o derives obstacle o's orientation, velocity, shape and responsibility flag from its
// raw prediction (`enrich`, optional) and packs that obstacle's rows of the tile -- a sample of the tile needs nothing of
// another obstacle, so the two steps meet at a block barrier instead of at a kernel boundary.
template <typename R>
__global__ void __launch_bounds__(128) cycle_head_kernel(const DevStep* __restrict__ steps, const __grid_constant__ DevParams pr,
                                                         const PackArgs* __restrict__ pack, const EnrichArgs* __restrict__ enrich, int n_prof) {
    extern __shared__ double sm[];
    if ((int)blockIdx.x < n_prof) {
        const DevStep& st = steps[0];
        profile_body(st, pr, st.nskip, sm, blockIdx.x);
        return;
    }
    const int o = blockIdx.x - n_prof;
    const PackArgs a = *pack;
    if (o >= a.O) return;
    if (enrich) {
        const EnrichArgs e = *enrich;
        enrich_body(e.O, e.Tp, e.pred_len, e.pos, e.raw_len, e.raw_wid, e.ori0, e.vel0, e.dt, e.margin_l, e.margin_w, e.ego_x, e.ego_y, e.ego_ori,
                    e.wrap_view_cone != 0, e.yaw, e.vel, e.length, e.width, e.resp, sm, o);
        __syncthreads();  // this block's orientation / velocity rows and shape are what its own samples read
    }
    for (int t = threadIdx.x; t < a.Tp; t += blockDim.x) pack_sample<R>(a, o * a.Tp + t);
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

// principle costs of one trajectory from per-obstacle maxima (risk_costs.py:128-255); one warp
__global__ void principles_kernel(int n, const double* __restrict__ v, const int* __restrict__ resp, double bd, double eps,
                                  double scale, bool relaxed_gate, double* __restrict__ out) {
    const int lane = threadIdx.x;
    double sum_e = 0, sum_o = 0, sum_abs = 0, rs = 0, mm = -INFINITY;
    for (int o = lane; o < n; o += 32) {
        const double re = v[o], ro = v[n + o], he = v[2 * n + o], ho = v[3 * n + o];
        sum_e += re;
        sum_o += ro;
        sum_abs += fabs(re - ro);
        rs -= (double)resp[o] * ro;
        mm = fmax(mm, he * ((re < eps) != relaxed_gate ? 1.0 : 0.0));  // risk_costs.py:205-206 (Q3)
        mm = fmax(mm, ho * ((ro < eps) != relaxed_gate ? 1.0 : 0.0));
    }
    for (int k = 16; k > 0; k >>= 1) {
        sum_e += __shfl_xor_sync(kFull, sum_e, k);
        sum_o += __shfl_xor_sync(kFull, sum_o, k);
        sum_abs += __shfl_xor_sync(kFull, sum_abs, k);
        rs += __shfl_xor_sync(kFull, rs, k);
        mm = fmax(mm, __shfl_xor_sync(kFull, mm, k));
    }
    if (lane == 0) {
        if (n == 0) {  // risk_costs.py:145-146,173-174,199-200,222-223
            for (int i = 0; i < 5; ++i) out[i] = 0.0;
        } else {
            out[0] = (sum_e + sum_o + bd) / (double)(n * 2);
            out[1] = sum_abs / (double)n;
            out[2] = pow(fmax(mm, bd), scale);
            out[3] = rs;
            out[4] = sum_e + bd;
        }
    }
}

// the winner of the last step, found on the device: [BestRec | int T | pad | fields[ETP_NUM_FIELDS][Tmax]] in one buffer
__global__ void best_trajectory_kernel(DevStep st, const BestRec* __restrict__ best, unsigned char* __restrict__ combo) {
    best_trajectory_body(st, *best, combo);
}

// the same for step record `scn` of a device array (scenario sweep / planning cycle)
__global__ void best_trajectory_steps_kernel(const DevStep* __restrict__ steps, const BestRec* __restrict__ best,
                                             unsigned char* __restrict__ combo) {
    best_trajectory_body(steps[0], *best, combo);
}

// lexicographic minimum of shard-local best records (one thread: a handful of 40-byte records), the device twin of
// etp_combine_best
__global__ void combine_best_kernel(const BestRec* __restrict__ recs, int n, int consecutive, BestRec* __restrict__ out) {
    if (threadIdx.x != 0) return;
    int win = -1, before = 0, scored = 0;
    for (int i = 0; i < n; ++i) {
        const BestRec b = recs[i];
        if (b.index >= 0 && (win < 0 || b.key_hi < recs[win].key_hi || (b.key_hi == recs[win].key_hi && b.key_lo < recs[win].key_lo))) {
            win = i;
            before = scored;
        }
        scored += b.n_scored;
    }
    BestRec r = recs[win < 0 ? 0 : win];
    r.n_scored = scored;
    if (win < 0) r.index = -1;
    else r.list_index = (consecutive ? before : 0) + r.list_index;
    *out = r;
}

cudaError_t launch_combine_best(const BestRec* recs, int n, int consecutive, BestRec* out, cudaStream_t stream) {
    combine_best_kernel<<<1, 32, 0, stream>>>(recs, n, consecutive, out);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// host-side launchers (called from etp_abi.cpp)
// ---------------------------------------------------------------------------------------------
cudaError_t launch_pack_obstacles(int precision, int O, int Tp, const int* pred_len, const double* pos, const double* cov,
                                  const double* yaw, const double* vel, const double* length, const double* width,
                                  const double* mass, const int* prot, const int* resp, double veh_m, double ox, double oy,
                                  void* tile,
                                  double* oconst, cudaStream_t stream) {
    const int n = O * Tp;
    if (n <= 0) return cudaSuccess;
    const int blocks = (n + 127) / 128;
    const PackArgs a{O, Tp, pred_len, pos, cov, yaw, vel, length, width, mass, prot, resp, veh_m, ox, oy, tile, oconst};
    if (precision == ETP_PRECISION_FP64) pack_obstacles_kernel<double><<<blocks, 128, 0, stream>>>(a);
    else pack_obstacles_kernel<float><<<blocks, 128, 0, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_pack_obstacles_batch(int precision, const PackArgs* args, int n_scn, int max_items, cudaStream_t stream) {
    if (n_scn <= 0 || max_items <= 0) return cudaSuccess;
    const dim3 grid((max_items + 127) / 128, n_scn);
    if (precision == ETP_PRECISION_FP64) pack_obstacles_batch_kernel<double><<<grid, 128, 0, stream>>>(args);
    else pack_obstacles_batch_kernel<float><<<grid, 128, 0, stream>>>(args);
    return cudaGetLastError();
}

cudaError_t launch_profiles(const DevStep& st, const DevParams& pr, int* nskip, cudaStream_t stream) {
    const int n_prof = st.p_end - st.p_begin;
    if (n_prof <= 0) return cudaSuccess;
    const size_t smem = sizeof(double) * 7 * st.Tmax;
    profile_kernel<<<n_prof, 128, smem, stream>>>(st, pr, nskip);
    return cudaGetLastError();
}

cudaError_t launch_profiles_batch(const DevStep* steps, const DevParams& pr, int n_scn, int max_prof, int max_T, cudaStream_t stream) {
    if (n_scn <= 0 || max_prof <= 0) return cudaSuccess;
    profile_batch_kernel<<<dim3(max_prof, n_scn), 128, sizeof(double) * 7 * max_T, stream>>>(steps, pr);
    return cudaGetLastError();
}

cudaError_t launch_enrich(int O, int Tp, const int* pred_len, const double* pos, const double* raw_len, const double* raw_wid,
                          const double* ori0, const double* vel0, double dt, double margin_l, double margin_w, double ego_x,
                          double ego_y, double ego_ori, bool wrap_view_cone, double* yaw, double* vel, double* length, double* width, int* resp,
                          cudaStream_t stream) {
    if (O <= 0) return cudaSuccess;
    enrich_kernel<<<O, 128, sizeof(double) * 3 * Tp, stream>>>(O, Tp, pred_len, pos, raw_len, raw_wid, ori0, vel0, dt, margin_l, margin_w,
                                                              ego_x, ego_y, ego_ori, wrap_view_cone, yaw, vel, length, width, resp);
    return cudaGetLastError();
}

cudaError_t launch_principles(int n, const double* vals, const int* resp, double bd, double eps, double scale, bool relaxed_gate,
                              double* out5, cudaStream_t stream) {
    principles_kernel<<<1, 32, 0, stream>>>(n, vals, resp, bd, eps, scale, relaxed_gate, out5);
    return cudaGetLastError();
}

cudaError_t launch_best_trajectory(const DevStep& st, const BestRec* best, unsigned char* combo, cudaStream_t stream) {
    best_trajectory_kernel<<<1, 64, 0, stream>>>(st, best, combo);
    return cudaGetLastError();
}

cudaError_t launch_enrich_args(const EnrichArgs* args, int O, int Tp, cudaStream_t stream) {
    if (O <= 0) return cudaSuccess;
    enrich_args_kernel<<<O, 128, sizeof(double) * 3 * Tp, stream>>>(args);
    return cudaGetLastError();
}

cudaError_t launch_cycle_head(int precision, const DevStep* steps, const DevParams& pr, const PackArgs* pack, const EnrichArgs* enrich,
                              int n_prof, int O, int Tp, int Tmax, cudaStream_t stream) {
    const int ma = 7 * Tmax, mb = 3 * Tp;
    const size_t smem = sizeof(double) * (size_t)(ma > mb ? ma : mb);
    if (precision == ETP_PRECISION_FP64) cycle_head_kernel<double><<<n_prof + O, 128, smem, stream>>>(steps, pr, pack, enrich, n_prof);
    else cycle_head_kernel<float><<<n_prof + O, 128, smem, stream>>>(steps, pr, pack, enrich, n_prof);
    return cudaGetLastError();
}

cudaError_t launch_best_trajectory_steps(const DevStep* steps, const BestRec* best, unsigned char* combo, cudaStream_t stream) {
    best_trajectory_steps_kernel<<<1, 64, 0, stream>>>(steps, best, combo);
    return cudaGetLastError();
}

cudaError_t launch_trajectory(const DevStep& st, int c_dense, double* fields, int* n_out, cudaStream_t stream) {
    trajectory_kernel<<<1, 64, 0, stream>>>(st, c_dense, fields, n_out);
    return cudaGetLastError();
}

}  // namespace etp
