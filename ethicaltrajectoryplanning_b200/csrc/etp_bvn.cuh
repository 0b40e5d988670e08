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

__device__ __forceinline__ double phi_d(double x) { return 0.5 * erfc(-x * 0.70710678118654752440); }

__constant__ double kGLw[3][10] = {
    {0.1713244923791705, 0.3607615730481384, 0.4679139345726904},
    {0.04717533638651177, 0.1069393259953183, 0.1600783285433464, 0.2031674267230659, 0.2334925365383547,
     0.2491470458134029},
    {0.01761400713915212, 0.04060142980038694, 0.06267204833410906, 0.08327674157670475, 0.1019301198172404,
     0.1181945319615184, 0.1316886384491766, 0.1420961093183821, 0.1491729864726037, 0.1527533871307259}};
__constant__ double kGLx[3][10] = {
    {0.9324695142031522, 0.6612093864662647, 0.2386191860831970},
    {0.9815606342467191, 0.9041172563704750, 0.7699026741943050, 0.5873179542866171, 0.3678314989981802,
     0.1252334085114692},
    {0.9931285991850949, 0.9639719272779138, 0.9122344282513259, 0.8391169718222188, 0.7463319064601508,
     0.6360536807265150, 0.5108670019508271, 0.3737060887154196, 0.2277858511416451, 0.07652652113349733}};

// Genz BVU: P(X > sh, Y > sk) for a standard bivariate normal with correlation r.
__device__ double bvu_d(double sh, double sk, double r) {
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

// Phi(b) - Phi(a), a <= b, from the tail side (no cancellation)
__device__ __forceinline__ float dphi_f(float a, float b) {
    const float is2 = 0.70710678118654752440f;
    if (a >= 0.0f) return 0.5f * (erfcf(a * is2) - erfcf(b * is2));
    if (b <= 0.0f) return 0.5f * (erfcf(-b * is2) - erfcf(-a * is2));
    return 0.5f * (erff(b * is2) + erff(-a * is2));
}

__device__ __forceinline__ float corner_sum_f(float xl, float xu, float yl, float yu, float s, float q) {
    // q = log2(e) / (1 - s^2); E = exp2(q * (s*h*k - (h^2+k^2)/2))
    const float hxu = 0.5f * xu * xu, hxl = 0.5f * xl * xl, hyu = 0.5f * yu * yu, hyl = 0.5f * yl * yl;
    const float e_uu = exp2f(q * (s * (xu * yu) - (hxu + hyu)));
    const float e_lu = exp2f(q * (s * (xl * yu) - (hxl + hyu)));
    const float e_ul = exp2f(q * (s * (xu * yl) - (hxu + hyl)));
    const float e_ll = exp2f(q * (s * (xl * yl) - (hxl + hyl)));
    return (e_uu - e_lu) - (e_ul - e_ll);
}

// fp32 rectangle probability, see file header.  `asr` = asin(r).
__device__ float rect_prob(float xl, float xu, float yl, float yu, float r) {
    const float ar = fabsf(r);
    if (ar > 0.55f) return (float)rect_prob((double)xl, (double)xu, (double)yl, (double)yu, (double)r);
    const float marg = dphi_f(xl, xu) * dphi_f(yl, yu);
    if (r == 0.0f) return marg;
    const float asr = asinf(r);
    const float LOG2E = 1.4426950408889634f;
    float acc = 0.0f;
    if (ar <= 0.35f) {
        const float gx[2] = {0.3399810435848563f, 0.8611363115940526f};
        const float gw[2] = {0.6521451548625461f, 0.3478548451374538f};
#pragma unroll
        for (int i = 0; i < 2; ++i) {
#pragma unroll
            for (int sg = -1; sg <= 1; sg += 2) {
                const float s = sinf(asr * (0.5f + 0.5f * sg * gx[i]));
                const float q = LOG2E / (1.0f - s * s);
                acc += gw[i] * corner_sum_f(xl, xu, yl, yu, s, q);
            }
        }
    } else {
        const float gx[3] = {0.2386191860831969f, 0.6612093864662645f, 0.9324695142031521f};
        const float gw[3] = {0.4679139345726910f, 0.3607615730481386f, 0.1713244923791704f};
#pragma unroll
        for (int i = 0; i < 3; ++i) {
#pragma unroll
            for (int sg = -1; sg <= 1; sg += 2) {
                const float s = sinf(asr * (0.5f + 0.5f * sg * gx[i]));
                const float q = LOG2E / (1.0f - s * s);
                acc += gw[i] * corner_sum_f(xl, xu, yl, yu, s, q);
            }
        }
    }
    return marg + acc * asr * 0.07957747154594767f;  // 1 / (4 pi)
}

}  // namespace etp
