// gprfit.hpp — per-factor hyper-parameter fit of the GPR depth factor (host side of the boundary).
//
// Replaces GPR::fit (include/GPR.hpp:350-387) and its objective GPRHyperLoss::Evaluate (GPR.hpp:154-174):
// (sigma, l) of the RBF kernel are chosen by minimising the negative log marginal likelihood of the
// neighbours' depths, starting from (init_sigma, init_l), at most 15 iterations (GPR.hpp:360).  The
// reference hands the objective to ceres::GradientProblemSolver (L-BFGS + Wolfe line search, a third-party
// dependency that is not vendored and not pinned); here the two-parameter problem is solved by a dense
// BFGS with a backtracking Armijo search.  What is reproduced exactly is the OBJECTIVE:
//   K = sigma^2 exp(-D / (2 l^2)) + sigma_n I,  alpha = K^-1 y (LLT),
//   nlml = 0.5 (y^T alpha + 2 sum log L_ii + n log 2 pi)                                  (GPR.hpp:163)
// and, selectable, the gradient in two flavours:
//   * kGradAsCoded — the expressions of GPR.hpp:166-171,218-222 as written: dK/dsigma = 2 sigma Kff with Kff
//     the FULL kernel matrix (sigma^2 and the noise term included) and dK/dl = (Kff * Dist) / l^3 with a
//     MATRIX product (Eigen operator* on two MatrixXd) where the derivative is element-wise;
//   * kGradAnalytic — the derivative of the objective (dK/dsigma = 2 sigma E, dK/dl = sigma^2 E o D / l^3,
//     E = exp(-D / 2 l^2)), which the finite-difference test checks.
// The iterate path of Ceres' L-BFGS cannot be reproduced without Ceres: a21 stays "parity-unpinned" on the
// path, pinned on the objective (tests/test_gpr_fit.py).
#pragma once
#include <cmath>
#include <vector>

namespace stl {

enum GprGrad { kGradAnalytic = 0, kGradAsCoded = 1 };

// squared pixel distances (self_pdist, GPR.hpp:41-54): D[i][j] = |x_i - x_j|^2, zero diagonal
inline void gpr_self_pdist(const double *x /*[n][2]*/, int n, std::vector<double> &D) {
    D.assign((size_t)n * n, 0.0);
    for (int r = 0; r + 1 < n; ++r)
        for (int c = r + 1; c < n; ++c) {
            const double dx = x[r * 2] - x[c * 2], dy = x[r * 2 + 1] - x[c * 2 + 1];
            const double d = dx * dx + dy * dy;
            D[(size_t)r * n + c] = d;
            D[(size_t)c * n + r] = d;
        }
}

// Objective and gradient at (sigma, l).  Returns false when the Cholesky factorisation fails
// (GPR.hpp:159-160: Evaluate returns false).  grad may be null.
inline bool gpr_nlml(const std::vector<double> &D, const double *y, int n, double sigma_noise, double sigma, double l, int flavour,
                     double *cost, double *grad) {
    std::vector<double> K((size_t)n * n), L((size_t)n * n, 0.0), alpha(n);
    const double s2 = sigma * sigma, coef = -0.5 / (l * l);
    for (size_t i = 0; i < (size_t)n * n; ++i) K[i] = s2 * std::exp(coef * D[i]);
    for (int i = 0; i < n; ++i) K[(size_t)i * n + i] += sigma_noise;
    // unblocked lower Cholesky (Eigen::LLT for small n)
    for (int k = 0; k < n; ++k) {
        double x = K[(size_t)k * n + k];
        for (int j = 0; j < k; ++j) x -= L[(size_t)k * n + j] * L[(size_t)k * n + j];
        if (!(x > 0.0)) return false;
        x = std::sqrt(x);
        L[(size_t)k * n + k] = x;
        for (int i = k + 1; i < n; ++i) {
            double v = K[(size_t)i * n + k];
            for (int j = 0; j < k; ++j) v -= L[(size_t)i * n + j] * L[(size_t)k * n + j];
            L[(size_t)i * n + k] = v / x;
        }
    }
    auto solve = [&](double *b) {  // b <- K^-1 b
        for (int i = 0; i < n; ++i) {
            double v = b[i];
            for (int j = 0; j < i; ++j) v -= L[(size_t)i * n + j] * b[j];
            b[i] = v / L[(size_t)i * n + i];
        }
        for (int i = n - 1; i >= 0; --i) {
            double v = b[i];
            for (int j = i + 1; j < n; ++j) v -= L[(size_t)j * n + i] * b[j];
            b[i] = v / L[(size_t)i * n + i];
        }
    };
    for (int i = 0; i < n; ++i) alpha[i] = y[i];
    solve(alpha.data());
    double ya = 0.0, logdet = 0.0;
    for (int i = 0; i < n; ++i) { ya += y[i] * alpha[i]; logdet += std::log(L[(size_t)i * n + i]); }
    const double two_pi = 6.283185307179586476925286766559;
    if (cost) *cost = 0.5 * (ya + 2.0 * logdet + n * std::log(two_pi));
    if (!grad) return true;
    // Kinv (GPR.hpp:197-200), inner = alpha alpha^T - Kinv, gradient_i = -0.5 tr(inner * dK_i)
    std::vector<double> Kinv((size_t)n * n, 0.0), col(n);
    for (int c = 0; c < n; ++c) {
        for (int i = 0; i < n; ++i) col[i] = i == c ? 1.0 : 0.0;
        solve(col.data());
        for (int i = 0; i < n; ++i) Kinv[(size_t)i * n + c] = col[i];
    }
    std::vector<double> dKs((size_t)n * n), dKl((size_t)n * n);
    const double inv_l3 = 1.0 / (l * l * l);
    if (flavour == kGradAsCoded) {
        for (size_t i = 0; i < (size_t)n * n; ++i) dKs[i] = 2.0 * sigma * K[i];
        for (int i = 0; i < n; ++i)                     // (Kff * Dist) * inv_l3, matrix product as written
            for (int j = 0; j < n; ++j) {
                double s = 0.0;
                for (int k = 0; k < n; ++k) s += K[(size_t)i * n + k] * D[(size_t)k * n + j];
                dKl[(size_t)i * n + j] = s * inv_l3;
            }
    } else {
        for (size_t i = 0; i < (size_t)n * n; ++i) {
            const double e = std::exp(coef * D[i]);
            dKs[i] = 2.0 * sigma * e;
            dKl[i] = s2 * e * D[i] * inv_l3;
        }
    }
    double g0 = 0.0, g1 = 0.0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            const double inner = alpha[i] * alpha[j] - Kinv[(size_t)i * n + j];  // entry (i, j)
            g0 += inner * dKs[(size_t)j * n + i];                                // tr(inner * dK) = sum_ij inner_ij dK_ji
            g1 += inner * dKl[(size_t)j * n + i];
        }
    grad[0] = -0.5 * g0;
    grad[1] = -0.5 * g1;
    return true;
}

struct GprFitResult {
    double sigma, l, cost0, cost;
    int iterations, evaluations;
    bool ok;  // false: the objective could not even be evaluated at the start (hyper-parameters stay as given)
};

// At most max_iter quasi-Newton iterations from (sigma0, l0) (options.max_num_iterations = 15, GPR.hpp:360).
inline GprFitResult gpr_fit(const double *x /*[n][2]*/, const double *y, int n, double sigma_noise, double sigma0, double l0, int max_iter,
                            int flavour) {
    GprFitResult r{sigma0, l0, 0.0, 0.0, 0, 0, false};
    if (n <= 0) return r;
    std::vector<double> D;
    gpr_self_pdist(x, n, D);
    double th[2] = {sigma0, l0}, f, g[2];
    r.evaluations = 1;
    if (!gpr_nlml(D, y, n, sigma_noise, th[0], th[1], flavour, &f, g)) return r;
    r.ok = true;
    r.cost0 = r.cost = f;
    double H[4] = {1.0, 0.0, 0.0, 1.0};  // inverse Hessian approximation
    const double gtol = 1e-10, ftol = 1e-6;  // ceres: gradient_tolerance 1e-10, function_tolerance 1e-6
    for (int it = 0; it < max_iter; ++it) {
        if (std::fmax(std::fabs(g[0]), std::fabs(g[1])) <= gtol) break;
        double d[2] = {-(H[0] * g[0] + H[1] * g[1]), -(H[2] * g[0] + H[3] * g[1])};
        double slope = d[0] * g[0] + d[1] * g[1];
        if (!(slope < 0.0)) { H[0] = H[3] = 1.0; H[1] = H[2] = 0.0; d[0] = -g[0]; d[1] = -g[1]; slope = -(g[0] * g[0] + g[1] * g[1]); }
        // first step: scaled like ceres' initial step (1 / |g|_inf), afterwards the quasi-Newton step
        double step = it == 0 ? 1.0 / std::fmax(std::fabs(g[0]), std::fabs(g[1])) : 1.0;
        double fn = f, gn[2] = {g[0], g[1]}, tn[2] = {th[0], th[1]};
        bool accepted = false;
        for (int ls = 0; ls < 20; ++ls) {  // ceres: max_num_line_search_step_size_iterations = 20
            tn[0] = th[0] + step * d[0];
            tn[1] = th[1] + step * d[1];
            ++r.evaluations;
            const bool ok = tn[0] != 0.0 && tn[1] != 0.0 && gpr_nlml(D, y, n, sigma_noise, tn[0], tn[1], flavour, &fn, gn);
            if (ok && std::isfinite(fn) && fn <= f + 1e-4 * step * slope) { accepted = true; break; }
            step *= 0.5;
        }
        if (!accepted) break;  // line search failed: keep the best point found (ceres returns the same way)
        const double s[2] = {tn[0] - th[0], tn[1] - th[1]}, yk[2] = {gn[0] - g[0], gn[1] - g[1]};
        const double sy = s[0] * yk[0] + s[1] * yk[1];
        if (sy > 1e-14) {  // BFGS update of the inverse Hessian
            const double rho = 1.0 / sy;
            const double Hy[2] = {H[0] * yk[0] + H[1] * yk[1], H[2] * yk[0] + H[3] * yk[1]};
            const double yHy = yk[0] * Hy[0] + yk[1] * Hy[1];
            for (int a = 0; a < 2; ++a)
                for (int b = 0; b < 2; ++b)
                    H[a * 2 + b] += rho * ((1.0 + rho * yHy) * s[a] * s[b] - Hy[a] * s[b] - s[a] * Hy[b]);
        }
        const double df = f - fn;
        th[0] = tn[0]; th[1] = tn[1]; f = fn; g[0] = gn[0]; g[1] = gn[1];
        r.iterations = it + 1;
        if (df <= ftol * std::fabs(f)) break;
    }
    r.sigma = th[0]; r.l = th[1]; r.cost = f;
    return r;
}

}  // namespace stl
