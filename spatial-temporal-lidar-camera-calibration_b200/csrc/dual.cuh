// dual.cuh — forward-mode dual numbers with 7 partials (host + device).
//
// Stands in for ceres::Jet<double, N> as instantiated by the reference's autodiff
// cost functions (AutoDiffCostFunction<...,7>, DynamicAutoDiffCostFunction<...,6>:
// include/IBACalib2.hpp:207,594,637) and g2o's G2O_MAKE_AUTO_AD_FUNCTIONS
// (include/IBACalib.hpp:154).  The 7 directions are the raw parameters
// [omega(3), upsilon(3), s] (VertexSim3::oplusImpl is plain addition,
// include/g2o_tools.h:21-24).  Product/quotient rules are written as in Jet so that the
// rounding of the partials matches an autodiff evaluation.
#pragma once
#include <cmath>

#ifdef __CUDACC__
#define STL_HD __host__ __device__ __forceinline__
#else
#define STL_HD inline
#endif

namespace stl {

struct D7 {
    double a;
    double v[7];
};

STL_HD D7 d7_const(double x) { D7 r; r.a = x; for (int i = 0; i < 7; ++i) r.v[i] = 0.0; return r; }
STL_HD D7 d7_var(double x, int k) { D7 r = d7_const(x); r.v[k] = 1.0; return r; }
STL_HD D7 operator+(const D7 &f, const D7 &g) { D7 r; r.a = f.a + g.a; for (int i = 0; i < 7; ++i) r.v[i] = f.v[i] + g.v[i]; return r; }
STL_HD D7 operator-(const D7 &f, const D7 &g) { D7 r; r.a = f.a - g.a; for (int i = 0; i < 7; ++i) r.v[i] = f.v[i] - g.v[i]; return r; }
STL_HD D7 operator-(const D7 &f) { D7 r; r.a = -f.a; for (int i = 0; i < 7; ++i) r.v[i] = -f.v[i]; return r; }
STL_HD D7 operator*(const D7 &f, const D7 &g) { D7 r; r.a = f.a * g.a; for (int i = 0; i < 7; ++i) r.v[i] = f.a * g.v[i] + f.v[i] * g.a; return r; }
STL_HD D7 operator/(const D7 &f, const D7 &g) {
    D7 r;
    const double gi = 1.0 / g.a, q = f.a * gi;
    r.a = q;
    for (int i = 0; i < 7; ++i) r.v[i] = (f.v[i] - q * g.v[i]) * gi;
    return r;
}
STL_HD D7 operator*(const D7 &f, double c) { D7 r; r.a = f.a * c; for (int i = 0; i < 7; ++i) r.v[i] = f.v[i] * c; return r; }
STL_HD D7 operator+(const D7 &f, double c) { D7 r = f; r.a = f.a + c; return r; }
STL_HD D7 operator-(const D7 &f, double c) { D7 r = f; r.a = f.a - c; return r; }

// host-only transcendental rules (Sim3Exp is evaluated on the host, once per candidate)
inline D7 d7_sqrt(const D7 &f) { D7 r; r.a = std::sqrt(f.a); const double t = 1.0 / (2.0 * r.a); for (int i = 0; i < 7; ++i) r.v[i] = t * f.v[i]; return r; }
inline D7 d7_cos(const D7 &f) { D7 r; r.a = std::cos(f.a); const double t = -std::sin(f.a); for (int i = 0; i < 7; ++i) r.v[i] = t * f.v[i]; return r; }
inline D7 d7_sin(const D7 &f) { D7 r; r.a = std::sin(f.a); const double t = std::cos(f.a); for (int i = 0; i < 7; ++i) r.v[i] = t * f.v[i]; return r; }
inline D7 d7_pow(const D7 &f, double p) { D7 r; r.a = std::pow(f.a, p); const double t = p * std::pow(f.a, p - 1.0); for (int i = 0; i < 7; ++i) r.v[i] = t * f.v[i]; return r; }

// Sim3Exp / SE3Exp on duals (g2o_tools.h:106-183): x[0..5] -> R (9), t (3)
inline void d7_se3_exp(const D7 *x, D7 *R, D7 *t) {
    const D7 theta = d7_sqrt((x[0] * x[0] + x[1] * x[1]) + x[2] * x[2]);
    const D7 z = d7_const(0.0);
    const D7 O[9] = {z, -x[2], x[1], x[2], z, -x[0], -x[1], x[0], z};
    D7 O2[9], V[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) O2[i * 3 + j] = (O[i * 3] * O[j] + O[i * 3 + 1] * O[3 + j]) + O[i * 3 + 2] * O[6 + j];
    if (theta.a < 1e-4) {
        const D7 half = d7_const(0.5), sixth = d7_const(1.0) / d7_const(6.0);
        for (int i = 0; i < 9; ++i) {
            const D7 I = d7_const((i % 4 == 0) ? 1.0 : 0.0);
            R[i] = (I + O[i]) + half * O2[i];
            V[i] = (I + half * O[i]) + sixth * O2[i];
        }
    } else {
        const D7 costh = d7_cos(theta), sinth = d7_sin(theta);
        const D7 invth2 = d7_pow(theta, -2.0), invth3 = d7_pow(theta, -3.0);
        const D7 a = sinth / theta, b = (d7_const(1.0) - costh) * invth2, c = (theta - sinth) * invth3;
        for (int i = 0; i < 9; ++i) {
            const D7 I = d7_const((i % 4 == 0) ? 1.0 : 0.0);
            R[i] = (I + a * O[i]) + b * O2[i];
            V[i] = (I + b * O[i]) + c * O2[i];
        }
    }
    for (int i = 0; i < 3; ++i) t[i] = (V[i * 3] * x[3] + V[i * 3 + 1] * x[4]) + V[i * 3 + 2] * x[5];
}

// What the linearisation kernel needs per parameter vector x.
struct LmCand {
    D7 R[9], t[3];      // Sim3Exp(x)            (IBA_PlaneFactor, IBACalib2.hpp:157)
    D7 Rlc[9], tlc[3];  // SE3Exp(-x[0:6])       (Point2Point/Point2Plane, IBACalib2.hpp:573-577,614-619)
    D7 s;               // x[6]
};

inline void make_lm_candidate(const double *x, LmCand *c) {
    D7 xd[7], xn[6];
    for (int i = 0; i < 7; ++i) xd[i] = d7_var(x[i], i);
    for (int i = 0; i < 6; ++i) xn[i] = -xd[i];
    d7_se3_exp(xd, c->R, c->t);
    d7_se3_exp(xn, c->Rlc, c->tlc);
    c->s = xd[6];
}

}  // namespace stl
