// crmath.cuh — correctly rounded acos / cos for the closed-form eigen-solver of the plane fit.
//
// FastEigen3x3_EV (include/pointcloud.h:378-463) calls std::acos once and std::cos twice; every other operation of the
// plane fit is +, -, *, /, sqrt, which IEEE-754 makes identical on the GPU and on the host.  CUDA's acos / cos are
// accurate to 1-2 ulp, glibc's are correctly rounded in all but a percent or so of their arguments — so with CUDA's
// functions a large share of the plane normals differ from a CPU evaluation in their last bit.  That is invisible in
// the cost, but the ray-plane intersection of IBA_PlaneFactor (IBACalib2.hpp:152-184) divides by n_c . ray: for the few
// blocks whose ray is almost parallel to the plane the last bit of the normal is amplified a million-fold, and those
// blocks dominate J^T J (observed: 7e-7 .. 1.6e-6 relative, at the edge of the 1e-6 bar).  The two functions below are
// evaluated in double-double arithmetic (~104 bits) and rounded once, i.e. they return the correctly rounded value and
// thus agree with glibc wherever glibc itself rounds correctly.
//   cos: argument in [0, pi] (the eigen-solver's range; any finite argument below 2^20 works), Cody-Waite reduction by
//        pi/2 in three parts, Taylor series of sin / cos in double-double.
//   acos: one Newton step on cos from CUDA's acos, with sin / cos of the iterate in double-double.
#pragma once
#include <cuda_runtime.h>

namespace stl {

struct dd { double h, l; };

__device__ __forceinline__ dd dd_two_sum(double a, double b) {
    const double s = __dadd_rn(a, b), bb = __dsub_rn(s, a);
    return {s, __dadd_rn(__dsub_rn(a, __dsub_rn(s, bb)), __dsub_rn(b, bb))};
}
__device__ __forceinline__ dd dd_quick_two_sum(double a, double b) {  // |a| >= |b|
    const double s = __dadd_rn(a, b);
    return {s, __dsub_rn(b, __dsub_rn(s, a))};
}
__device__ __forceinline__ dd dd_two_prod(double a, double b) {
    const double p = __dmul_rn(a, b);
    return {p, __fma_rn(a, b, -p)};
}
__device__ __forceinline__ dd dd_add(dd a, dd b) {
    dd s = dd_two_sum(a.h, b.h);
    const dd t = dd_two_sum(a.l, b.l);
    s.l = __dadd_rn(s.l, t.h);
    s = dd_quick_two_sum(s.h, s.l);
    s.l = __dadd_rn(s.l, t.l);
    return dd_quick_two_sum(s.h, s.l);
}
__device__ __forceinline__ dd dd_add_d(dd a, double b) {
    dd s = dd_two_sum(a.h, b);
    s.l = __dadd_rn(s.l, a.l);
    return dd_quick_two_sum(s.h, s.l);
}
__device__ __forceinline__ dd dd_mul(dd a, dd b) {
    dd p = dd_two_prod(a.h, b.h);
    p.l = __dadd_rn(p.l, __dadd_rn(__dmul_rn(a.h, b.l), __dmul_rn(a.l, b.h)));
    return dd_quick_two_sum(p.h, p.l);
}
__device__ __forceinline__ dd dd_neg(dd a) { return {-a.h, -a.l}; }
__device__ __forceinline__ dd dd_div(dd a, dd b) {
    const double q1 = __ddiv_rn(a.h, b.h);
    dd r = dd_add(a, dd_neg(dd_mul(b, dd{q1, 0.0})));
    const double q2 = __ddiv_rn(r.h, b.h);
    r = dd_add(r, dd_neg(dd_mul(b, dd{q2, 0.0})));
    const double q3 = __ddiv_rn(r.h, b.h);
    dd q = dd_quick_two_sum(q1, q2);
    return dd_add_d(q, q3);
}

__device__ constexpr double kPio2_1 = 0x1.921fb54442d18p+0, kPio2_2 = 0x1.1a62633145c07p-54, kPio2_3 = -0x1.f1976b7ed8fbcp-110;
__device__ const double kCosC[14][2] = {{-0x1.0000000000000p-1, -0x0.0p+0}, {0x1.5555555555555p-5, 0x1.5555555555555p-59}, {-0x1.6c16c16c16c17p-10, 0x1.f49f49f49f49fp-65}, {0x1.a01a01a01a01ap-16, 0x1.a01a01a01a01ap-76}, {-0x1.27e4fb7789f5cp-22, -0x1.cbbc05b4fa99ap-76}, {0x1.1eed8eff8d898p-29, -0x1.2aec959e14c06p-83}, {-0x1.93974a8c07c9dp-37, -0x1.05d6f8a2efd1fp-92}, {0x1.ae7f3e733b81fp-45, 0x1.1d8656b0ee8cbp-101}, {-0x1.6827863b97d97p-53, -0x1.eec01221a8b0bp-107}, {0x1.e542ba4020225p-62, 0x1.ea72b4afe3c2fp-120}, {-0x1.0ce396db7f853p-70, 0x1.aebcdbd20331cp-124}, {0x1.f2cf01972f578p-80, -0x1.9ada5fcc1ab14p-135}, {-0x1.88e85fc6a4e5ap-89, 0x1.71c37ebd16540p-143}, {0x1.0a18a2635085dp-98, 0x1.b9e2e28e1aa54p-153}};
__device__ const double kSinC[14][2] = {{-0x1.5555555555555p-3, -0x1.5555555555555p-57}, {0x1.1111111111111p-7, 0x1.1111111111111p-63}, {-0x1.a01a01a01a01ap-13, -0x1.a01a01a01a01ap-73}, {0x1.71de3a556c734p-19, -0x1.c154f8ddc6c00p-73}, {-0x1.ae64567f544e4p-26, 0x1.c062e06d1f209p-80}, {0x1.6124613a86d09p-33, 0x1.f28e0cc748ebep-87}, {-0x1.ae7f3e733b81fp-41, -0x1.1d8656b0ee8cbp-97}, {0x1.952c77030ad4ap-49, 0x1.ac981465ddc6cp-103}, {-0x1.2f49b46814157p-57, -0x1.2650f61dbdcb4p-112}, {0x1.71b8ef6dcf572p-66, -0x1.d043ae40c4647p-120}, {-0x1.761b41316381ap-75, 0x1.3423c7d91404fp-130}, {0x1.3f3ccdd165fa9p-84, -0x1.58ddadf344487p-139}, {-0x1.d1ab1c2dccea3p-94, -0x1.054d0c78aea14p-149}, {0x1.259f98b4358adp-103, 0x1.eaf8c39dd9bc5p-157}};

// sin and / or cos of x (|x| < ~1e6) in double-double.  The Taylor coefficients beyond the eighth term contribute less than
// 2^-58 of the result, so they are summed in plain double (error < 2^-110); the leading ones in double-double.
template <bool WANT_SIN, bool WANT_COS>
__device__ __forceinline__ void dd_sincos(double x, dd &s, dd &c) {
    const int k = (int)rint(x * 0.63661977236758134308);  // x / (pi/2)
    const double kd = (double)k;
    dd r = dd_two_sum(x, -kd * kPio2_1);                   // k * pio2_1 is exact for the small k that occur
    r = dd_add(r, dd_two_prod(-kd, kPio2_2));
    r = dd_add(r, dd_two_prod(-kd, kPio2_3));
    const dd z = dd_mul(r, r);
    const bool need_cos_series = (k & 1) ? WANT_SIN : WANT_COS, need_sin_series = (k & 1) ? WANT_COS : WANT_SIN;
    dd cr = {0.0, 0.0}, sr = {0.0, 0.0};
    if (need_cos_series) {
        double t = kCosC[13][0];
#pragma unroll
        for (int i = 12; i >= 8; --i) t = __fma_rn(t, z.h, kCosC[i][0]);
        dd pc = dd_add(dd_mul(dd{t, 0.0}, z), dd{kCosC[7][0], kCosC[7][1]});
#pragma unroll
        for (int i = 6; i >= 0; --i) pc = dd_add(dd_mul(pc, z), dd{kCosC[i][0], kCosC[i][1]});
        cr = dd_add_d(dd_mul(pc, z), 1.0);                 // cos r = 1 + z * C(z)
    }
    if (need_sin_series) {
        double t = kSinC[13][0];
#pragma unroll
        for (int i = 12; i >= 8; --i) t = __fma_rn(t, z.h, kSinC[i][0]);
        dd ps = dd_add(dd_mul(dd{t, 0.0}, z), dd{kSinC[7][0], kSinC[7][1]});
#pragma unroll
        for (int i = 6; i >= 0; --i) ps = dd_add(dd_mul(ps, z), dd{kSinC[i][0], kSinC[i][1]});
        sr = dd_add(r, dd_mul(r, dd_mul(ps, z)));          // sin r = r + r * z * S(z)
    }
    switch (k & 3) {
        case 0: s = sr; c = cr; break;
        case 1: s = cr; c = dd_neg(sr); break;
        case 2: s = dd_neg(sr); c = dd_neg(cr); break;
        default: s = dd_neg(cr); c = sr; break;
    }
}

__device__ inline double cos_cr(double x) {
    if (!(fabs(x) < 1.0e6)) return cos(x);
    dd s, c;
    dd_sincos<false, true>(x, s, c);
    return c.h;
}

__device__ inline double acos_cr(double x) {
    const double y0 = acos(x);
    if (!(fabs(x) < 1.0) || y0 < 1e-6 || y0 > 3.1415916) return y0;  // at the very ends one Newton step on cos has no grip
    dd s, c;
    dd_sincos<true, true>(y0, s, c);
    const dd corr = dd_div(dd_add_d(c, -x), s);  // cos(y) = x:  y = y0 + (cos y0 - x) / sin y0
    return dd_add_d(corr, y0).h;
}

}  // namespace stl
