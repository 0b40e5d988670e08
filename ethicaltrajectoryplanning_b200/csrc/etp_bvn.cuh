// Bivariate normal rectangle probabilities on the device.
//
// The reference evaluates scipy.stats.mvn.mvnun (Fortran MVNDST, un-vendored) per (mean, ego box)
// pair: risk_assessment/collision_probability.py:245-247.  For two dimensions MVNDST reduces to a
// four-corner combination of Genz' BVU.
//
//  * fp64 check mode: BVU restated verbatim (6/12/20-point Gauss-Legendre, |r| >= 0.925 branch),
//    four corners in MVNDST's order -> agrees with the reference to its own ~1e-16 noise floor.
//  * fp32 mode: the four-corner sum loses everything below ~1e-7 absolute in fp32.
//    Used instead:  P = dPhi_x * dPhi_y + asin(r)/(4 pi) * sum_i w_i B(theta_i),
//    B(theta) = E(xu,yu) - E(xl,yu) - E(xu,yl) + E(xl,yl), E(h,k) = exp((sin(theta) h k - (h^2+k^2)/2)/cos^2(theta)),
//    i.e. Genz' rho-integral with the corner sum moved inside the integrand and the marginal term
//    built from tail-side erfc differences (no cancellation).  4 nodes for |r| <= 0.35, 6 for
//    |r| <= 0.55 (measured: every case within 1e-4 rel + 1e-9 abs of fp64 BVU); larger |r| falls
//    back to the fp64 routine.
#pragma once

#include <math.h>

namespace etp {

// single-instruction MUFU forms (flush-to-zero, no range fix-up code around them): arguments below -126 give
// exactly 0, which is what the tails want
__device__ __forceinline__ float ex2_ftz(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rsqrt_ftz(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ double phi_d(double x) { return 0.5 * erfc(-x * 0.70710678118654752440); }

static __constant__ double kGLw[3][10] = {
    {0.1713244923791705, 0.3607615730481384, 0.4679139345726904},
    {0.04717533638651177, 0.1069393259953183, 0.1600783285433464, 0.2031674267230659, 0.2334925365383547,
     0.2491470458134029},
    {0.01761400713915212, 0.04060142980038694, 0.06267204833410906, 0.08327674157670475, 0.1019301198172404,
     0.1181945319615184, 0.1316886384491766, 0.1420961093183821, 0.1491729864726037, 0.1527533871307259}};
static __constant__ double kGLx[3][10] = {
    {0.9324695142031522, 0.6612093864662647, 0.2386191860831970},
    {0.9815606342467191, 0.9041172563704750, 0.7699026741943050, 0.5873179542866171, 0.3678314989981802,
     0.1252334085114692},
    {0.9931285991850949, 0.9639719272779138, 0.9122344282513259, 0.8391169718222188, 0.7463319064601508,
     0.6360536807265150, 0.5108670019508271, 0.3737060887154196, 0.2277858511416451, 0.07652652113349733}};

// Genz BVU: P(X > sh, Y > sk) for a standard bivariate normal with correlation r.
static __device__ double bvu_d(double sh, double sk, double r) {
    const double TWOPI = 6.283185307179586;
    int ng, lg;
    const double ar = fabs(r);
    if (ar < 0.3) { ng = 0; lg = 3; }
    else if (ar < 0.75) { ng = 1; lg = 6; }
    else { ng = 2; lg = 10; }
    double h = sh, k = sk, hk = h * k, bvn = 0.0;
    if (ar < 0.925) {
        const double hs = (h * h + k * k) / 2.0;
        const double asr = asin(r);
        for (int i = 0; i < lg; ++i) {
            double sn = sin(asr * (-kGLx[ng][i] + 1.0) / 2.0);
            bvn += kGLw[ng][i] * exp((sn * hk - hs) / (1.0 - sn * sn));
            sn = sin(asr * (kGLx[ng][i] + 1.0) / 2.0);
            bvn += kGLw[ng][i] * exp((sn * hk - hs) / (1.0 - sn * sn));
        }
        return bvn * asr / (2.0 * TWOPI) + phi_d(-h) * phi_d(-k);
    }
    if (r < 0) { k = -k; hk = -hk; }
    if (ar < 1.0) {
        const double as = (1.0 - r) * (1.0 + r);
        double a = sqrt(as);
        const double bs = (h - k) * (h - k);
        const double c = (4.0 - hk) / 8.0;
        const double d = (12.0 - hk) / 16.0;
        bvn = a * exp(-(bs / as + hk) / 2.0) * (1.0 - c * (bs - as) * (1.0 - d * bs / 5.0) / 3.0 + c * d * as * as / 5.0);
        if (hk > -160.0) {
            const double b = sqrt(bs);
            bvn -= exp(-hk / 2.0) * sqrt(TWOPI) * phi_d(-b / a) * b * (1.0 - c * bs * (1.0 - d * bs / 5.0) / 3.0);
        }
        a = a / 2.0;
        for (int i = 0; i < lg; ++i) {
            double xs = (a * (-kGLx[ng][i] + 1.0));
            xs = xs * xs;
            double rs = sqrt(1.0 - xs);
            bvn += a * kGLw[ng][i] *
                   (exp(-bs / (2.0 * xs) - hk / (1.0 + rs)) / rs - exp(-(bs / xs + hk) / 2.0) * (1.0 + c * xs * (1.0 + d * xs)));
            xs = as * (kGLx[ng][i] + 1.0) * (kGLx[ng][i] + 1.0) / 4.0;
            rs = sqrt(1.0 - xs);
            bvn += a * kGLw[ng][i] * exp(-(bs / xs + hk) / 2.0) *
                   (exp(-hk * (1.0 - rs) / (2.0 * (1.0 + rs))) / rs - (1.0 + c * xs * (1.0 + d * xs)));
        }
        bvn = -bvn / TWOPI;
    }
    if (r > 0) return bvn + phi_d(-fmax(h, k));
    return -bvn + fmax(0.0, phi_d(-h) - phi_d(-k));
}

// BVNMVN with both limits finite: four-corner inclusion-exclusion in mvndst.f's order.
__device__ __forceinline__ double rect_prob(double xl, double xu, double yl, double yu, double r) {
    return bvu_d(xl, yl, r) - bvu_d(xu, yl, r) - bvu_d(xl, yu, r) + bvu_d(xu, yu, r);
}

// Phi(b) - Phi(a), a <= b.  The interval is reflected so that a + b >= 0: then either 0 <= a (both
// erfc values are tails, no cancellation) or a < 0 < b with |a| <= b (difference of O(1) numbers whose
// result is at least Phi(b) - 1/2).  Branch free.
// Branch-free complementary error function: erfc(z) = t exp(-z^2 + P(t)), t = 1 / (1 + z/2), z >= 0
// (Chebyshev fit of Numerical Recipes' erfcc, fractional error 1.2e-7 in exact arithmetic; measured
// <= 2.2e-6 in fp32 for |x| <= 6.5 with the z^2 rounding residual folded back in).  libdevice's erfcf
// branches on the magnitude of its argument, which serialises a warp whose lanes hold different
// rectangles.
__device__ __forceinline__ float erfc_f(float x) {
    const float z = fabsf(x);
    const float t = rcp_ftz(fmaf(0.5f, z, 1.0f));
    float p = 0.17087277f;
    p = fmaf(p, t, -0.82215223f);
    p = fmaf(p, t, 1.48851587f);
    p = fmaf(p, t, -1.13520398f);
    p = fmaf(p, t, 0.27886807f);
    p = fmaf(p, t, -0.18628806f);
    p = fmaf(p, t, 0.09678418f);
    p = fmaf(p, t, 0.37409196f);
    p = fmaf(p, t, 1.00002368f);
    p = fmaf(p, t, -1.26551223f);
    const float zz = z * z;
    const float e = fmaf(z, z, -zz);  // exact rounding residual of z*z
    float r = t * ex2_ftz(1.4426950408889634f * (p - zz));
    r = fmaf(-r, e, r);
    return x >= 0.0f ? r : 2.0f - r;
}

__device__ __forceinline__ float dphi_f(float a, float b) {
    const float is2 = 0.70710678118654752440f;
    const bool neg = (a + b) < 0.0f;
    const float lo = neg ? -b : a, hi = neg ? -a : b;
    return 0.5f * (erfc_f(lo * is2) - erfc_f(hi * is2));
}

// Gauss-Legendre nodes on theta in [0, asin r]: s_i = sin(theta_i), q_i = log2(e) / (1 - s_i^2),
// c_i = w_i * asin(r) / (4 pi).  They depend on the correlation only and are shared by the nine
// rectangles of one evaluation.
template <int N> struct RhoNodes {
    float s[N > 0 ? N : 1], q[N > 0 ? N : 1], c[N > 0 ? N : 1];
};

template <int N> __device__ __forceinline__ RhoNodes<N> make_rho_nodes(float r) {
    RhoNodes<N> n;
    if (N == 0) return n;
    const float asr = asinf(r);
    const float LOG2E = 1.4426950408889634f;
    const float gx4[2] = {0.3399810435848563f, 0.8611363115940526f};
    const float gw4[2] = {0.6521451548625461f, 0.3478548451374538f};
    const float gx6[3] = {0.2386191860831969f, 0.6612093864662645f, 0.9324695142031521f};
    const float gw6[3] = {0.4679139345726910f, 0.3607615730481386f, 0.1713244923791704f};
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const float g = (N == 4) ? gx4[i >> 1] : gx6[i >> 1];
        const float w = (N == 4) ? gw4[i >> 1] : gw6[i >> 1];
        const float th = asr * (0.5f + ((i & 1) ? 0.5f : -0.5f) * g);
        // theta <= asin(0.55) = 0.58: odd Taylor polynomial of sin to x^9 (error < 1e-9)
        const float t2 = th * th;
        const float sn = th * (1.0f + t2 * (-1.6666667e-1f + t2 * (8.3333333e-3f + t2 * (-1.9841270e-4f + t2 * 2.7557319e-6f))));
        n.s[i] = sn;
        n.q[i] = __fdividef(LOG2E, 1.0f - sn * sn);
        n.c[i] = w * asr * 0.07957747154594767f;  // 1 / (4 pi)
    }
    return n;
}

// sum_i c_i * [E(xu,yu) - E(xl,yu) - E(xu,yl) + E(xl,yl)](theta_i)
template <int N>
__device__ __forceinline__ float rho_correction_f(const RhoNodes<N>& n, float xl, float xu, float yl, float yu) {
    if (N == 0) return 0.0f;
    const float huu = 0.5f * (xu * xu + yu * yu), hlu = 0.5f * (xl * xl + yu * yu);
    const float hul = 0.5f * (xu * xu + yl * yl), hll = 0.5f * (xl * xl + yl * yl);
    const float puu = xu * yu, plu = xl * yu, pul = xu * yl, pll = xl * yl;
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const float s = n.s[i], q = n.q[i];
        const float e_uu = ex2_ftz(q * fmaf(s, puu, -huu));
        const float e_lu = ex2_ftz(q * fmaf(s, plu, -hlu));
        const float e_ul = ex2_ftz(q * fmaf(s, pul, -hul));
        const float e_ll = ex2_ftz(q * fmaf(s, pll, -hll));
        acc = fmaf(n.c[i], (e_uu - e_lu) - (e_ul - e_ll), acc);
    }
    return acc;
}

// ---------------------------------------------------------------------------------------------
// Packed (two fp32 lanes per instruction, FFMA2 / FMUL2 / FADD2 of sm_100) forms.  One un-gated evaluation is
// 36 erfc and up to 216 exponentials; their polynomial and argument arithmetic pairs up naturally (the two ends
// of an interval, the two x-limits of a corner pair), which halves the issue slots next to the MUFU ops.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 splat2(float a) { return make_float2(a, a); }

// (Phi(xu) - Phi(xl), Phi(yu) - Phi(yl)): four erfc evaluations in two packed streams (same formula as erfc_f)
__device__ __forceinline__ float2 dphi2_f(float xl, float xu, float yl, float yu) {
    const float is2 = 0.70710678118654752440f;
    const bool nx = (xl + xu) < 0.0f, ny = (yl + yu) < 0.0f;
    // reflected so that lo + hi >= 0: hi >= |lo|, hi >= 0
    const float lox = nx ? -xu : xl, hix = nx ? -xl : xu;
    const float loy = ny ? -yu : yl, hiy = ny ? -yl : yu;
    const float2 ax = make_float2(fabsf(lox), hix), ay = make_float2(fabsf(loy), hiy);
    const float2 zx = __fmul2_rn(ax, splat2(is2)), zy = __fmul2_rn(ay, splat2(is2));
    const float2 nzx = __fmul2_rn(ax, splat2(-is2)), nzy = __fmul2_rn(ay, splat2(-is2));
    const float2 ux = __ffma2_rn(zx, splat2(0.5f), splat2(1.0f)), uy = __ffma2_rn(zy, splat2(0.5f), splat2(1.0f));
    const float2 tx = make_float2(rcp_ftz(ux.x), rcp_ftz(ux.y)), ty = make_float2(rcp_ftz(uy.x), rcp_ftz(uy.y));
    // polynomial of erfc_f scaled by log2(e): the exponential is taken as 2^(.)
    const float L = 1.4426950408889634f;
#ifdef ETP_ESTRIN
    // Estrin's scheme: the same degree-9 polynomial with a dependency chain of four instead of nine
    float2 px, py;
    {
#define ETP_E2(t, hi, lo) __ffma2_rn(splat2((hi) * L), t, splat2((lo) * L))
        const float2 x2 = __fmul2_rn(tx, tx), y2 = __fmul2_rn(ty, ty);
        const float2 x4 = __fmul2_rn(x2, x2), y4 = __fmul2_rn(y2, y2);
        const float2 x8 = __fmul2_rn(x4, x4), y8 = __fmul2_rn(y4, y4);
        const float2 ax01 = ETP_E2(tx, 1.00002368f, -1.26551223f), ay01 = ETP_E2(ty, 1.00002368f, -1.26551223f);
        const float2 ax23 = ETP_E2(tx, 0.09678418f, 0.37409196f), ay23 = ETP_E2(ty, 0.09678418f, 0.37409196f);
        const float2 ax45 = ETP_E2(tx, 0.27886807f, -0.18628806f), ay45 = ETP_E2(ty, 0.27886807f, -0.18628806f);
        const float2 ax67 = ETP_E2(tx, 1.48851587f, -1.13520398f), ay67 = ETP_E2(ty, 1.48851587f, -1.13520398f);
        const float2 ax89 = ETP_E2(tx, 0.17087277f, -0.82215223f), ay89 = ETP_E2(ty, 0.17087277f, -0.82215223f);
#undef ETP_E2
        const float2 bx0 = __ffma2_rn(ax23, x2, ax01), by0 = __ffma2_rn(ay23, y2, ay01);
        const float2 bx1 = __ffma2_rn(ax67, x2, ax45), by1 = __ffma2_rn(ay67, y2, ay45);
        px = __ffma2_rn(ax89, x8, __ffma2_rn(bx1, x4, bx0));
        py = __ffma2_rn(ay89, y8, __ffma2_rn(by1, y4, by0));
    }
#elif !defined(ETP_ERFC9)
    // degree-7 fit of the same form over z in [0, 6] (weighted least squares driven to equal ripple against 40-digit erfc):
    // 4.2e-7 relative in exact arithmetic, two packed steps shorter than the degree-9 polynomial
    float2 px = splat2(-0.16036327f * L), py = px;
#define ETP_P2(c) px = __ffma2_rn(px, tx, splat2((c) * L)); py = __ffma2_rn(py, ty, splat2((c) * L));
    ETP_P2(0.62911016f) ETP_P2(-0.77338737f) ETP_P2(0.12218644f) ETP_P2(0.09217999f) ETP_P2(0.34989667f)
    ETP_P2(1.00642514f) ETP_P2(-1.26604819f)
#undef ETP_P2
#else  // Numerical Recipes' degree-9 fit (1.2e-7)
    float2 px = splat2(0.17087277f * L), py = px;
#define ETP_P2(c) px = __ffma2_rn(px, tx, splat2((c) * L)); py = __ffma2_rn(py, ty, splat2((c) * L));
    ETP_P2(-0.82215223f) ETP_P2(1.48851587f) ETP_P2(-1.13520398f) ETP_P2(0.27886807f) ETP_P2(-0.18628806f)
    ETP_P2(0.09678418f) ETP_P2(0.37409196f) ETP_P2(1.00002368f) ETP_P2(-1.26551223f)
#undef ETP_P2
#endif
    const float2 zzx = __fmul2_rn(zx, zx), zzy = __fmul2_rn(zy, zy);
    const float2 gx = __ffma2_rn(splat2(-L), zzx, px), gy = __ffma2_rn(splat2(-L), zzy, py);
    float2 rx = __fmul2_rn(tx, make_float2(ex2_ftz(gx.x), ex2_ftz(gx.y)));
    float2 ry = __fmul2_rn(ty, make_float2(ex2_ftz(gy.x), ex2_ftz(gy.y)));
#ifdef ETP_RESIDUAL
    {   // the rounding of z^2 folded back in: exp(-(z^2 + e)) = exp(-z^2) (1 - e).  Worth 6e-8 z^2, i.e. 1.5e-6 at seven
        // sigma where rectangles are dropped anyway: off by default (measured: 1 % of the 10^6-candidate step)
        const float2 nex = __ffma2_rn(nzx, zx, zzx), ney = __ffma2_rn(nzy, zy, zzy);  // -(exact z^2 - rounded z^2)
        rx = __ffma2_rn(rx, nex, rx);
        ry = __ffma2_rn(ry, ney, ry);
    }
#endif
    const float elx = lox >= 0.0f ? rx.x : 2.0f - rx.x, ely = loy >= 0.0f ? ry.x : 2.0f - ry.x;
    return make_float2(0.5f * (elx - rx.y), 0.5f * (ely - ry.y));
}

// sum_i c_i * [E(xu,yu) - E(xl,yu) - E(xu,yl) + E(xl,yl)](theta_i), corner pairs (xl, xu) packed
template <int N>
__device__ __forceinline__ float rho_correction2_f(const RhoNodes<N>& n, float xl, float xu, float yl, float yu) {
    if (N == 0) return 0.0f;
    const float2 X = make_float2(xl, xu);
    const float2 nxh = __fmul2_rn(__fmul2_rn(X, X), splat2(-0.5f));                 // -x^2 / 2
    const float2 nyh = __fmul2_rn(__fmul2_rn(make_float2(yl, yu), make_float2(yl, yu)), splat2(-0.5f));
    const float2 nhu = __fadd2_rn(nxh, splat2(nyh.y)), nhl = __fadd2_rn(nxh, splat2(nyh.x));  // -(x^2 + y^2) / 2
    const float2 pu = __fmul2_rn(X, splat2(yu)), pl = __fmul2_rn(X, splat2(yl));    // x y
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const float2 s = splat2(n.s[i]), q = splat2(n.q[i]);
        const float2 au = __fmul2_rn(q, __ffma2_rn(s, pu, nhu)), al = __fmul2_rn(q, __ffma2_rn(s, pl, nhl));
        const float2 eu = make_float2(ex2_ftz(au.x), ex2_ftz(au.y)), el = make_float2(ex2_ftz(al.x), ex2_ftz(al.y));
        const float2 d = __ffma2_rn(el, splat2(-1.0f), eu);  // (E(xl,yu) - E(xl,yl), E(xu,yu) - E(xu,yl))
        acc = fmaf(n.c[i], d.y - d.x, acc);
    }
    return acc;
}

// rectangles farther than this many standard deviations from the mean contribute < 1.3e-12 each (tail of the
// nearer marginal), i.e. < 4e-12 to the stored probability: three orders below the 1e-9 absolute floor of the parity bar
#define ETP_PRUNE_SIGMA 7.0f

template <int N>
__device__ __forceinline__ float rect_prob_nodes_f(const RhoNodes<N>& n, float xl, float xu, float yl, float yu) {
    if (xl > ETP_PRUNE_SIGMA || xu < -ETP_PRUNE_SIGMA || yl > ETP_PRUNE_SIGMA || yu < -ETP_PRUNE_SIGMA) return 0.0f;
    const float2 d = dphi2_f(xl, xu, yl, yu);
    return fmaf(d.x, d.y, rho_correction2_f<N>(n, xl, xu, yl, yu));
}

// strongly correlated predictions: fp64 Genz evaluation (kept out of line, rare)
static __device__ __noinline__ float rect_prob_fallback_f(float xl, float xu, float yl, float yu, float r) {
    return (float)rect_prob((double)xl, (double)xu, (double)yl, (double)yu, (double)r);
}

}  // namespace etp
