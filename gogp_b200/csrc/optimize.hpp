// Hyper-parameter optimisation loops that drive the device evaluation without leaving the
// process: the tutorial's two MLE drivers,
//   "adam"   infer.Adam{Rate}.Step in a loop, stop when every |g_i| < THRESHOLD
//            (tutorial/tutorial.go:156-168);
//   "lbfgs"  optimize.Minimize(Func, Grad) with MajorIterations = ITERS and
//            GradientThreshold = THRESHOLD (tutorial/tutorial.go:131-155).
// gonum (v0.9.3) and infergo (v1.2.2) are not vendored in the reference tree, so these are
// the published algorithms restated, not ports: Adam with bias correction (Kingma & Ba),
// L-BFGS with the two-loop recursion and a bisection line search on the weak Wolfe
// conditions (decrease 1e-4, curvature 0.9; first trial step 1/|g|, later 1).
//
// Host-only and templated on the objective so the loops are unit-tested on a CPU against
// analytic functions (tests/cpp_opt_test.cc); the product instantiates them over
// gogp_observe + gogp_gradient (capi.cu), with X and Y resident in HBM: per iteration only
// P parameters go down and P + 1 numbers come back.
//
// The objective is MAXIMISED (log marginal likelihood + log prior).  It comes as the reference's own
// pair: value(x, &f) (gp.GP.Observe + priors: N^3/3 flop) and grad(g) for the point value() was last
// called at (gp.GP.Gradient: another 2N^3/3).  The line search asks for the gradient only once a
// trial point has passed the sufficient-decrease test, so a rejected step costs a third of an
// accepted one.  value() returns false where the reference would panic (covariance not positive
// definite): the line search treats that as "step too long"; Adam stops there.
#pragma once
#include <cmath>
#include <cstdint>
#include <deque>
#include <vector>

namespace gogp {

struct OptSettings {
    int method = 0;  // 0 adam, 1 lbfgs
    int max_iters = 100;
    double threshold = 1e-6;
    double rate = 0.01;
    double beta1 = 0.9, beta2 = 0.999, eps = 1e-8;
    int history = 15;
};

struct OptResult {
    int iters = 0;      // major iterations (Adam steps / L-BFGS directions)
    int evals = 0;      // objective evaluations
    int grads = 0;      // gradient evaluations (<= evals)
    double f0 = 0.0;    // objective at the starting point
    double f = 0.0;     // objective at the returned point
    int converged = 0;  // 1: every |g_i| < threshold
    int failed = 0;     // 1: the starting point itself could not be evaluated
};

inline bool below_threshold(const std::vector<double>& g, double thr) {
    for (double v : g)
        if (!(std::fabs(v) < thr)) return false;
    return true;
}

// infer.Adam ascent.  On return x is the last point reached and f the objective there (the
// tutorial's "final log likelihood", tutorial/tutorial.go:171-175); a point that cannot be
// evaluated is rolled back to its predecessor.
template <class Value, class Grad>
OptResult adam_ascent(Value&& value, Grad&& grad, std::vector<double>& x, const OptSettings& s) {
    OptResult r;
    const size_t n = x.size();
    std::vector<double> g(n), m(n, 0.0), v(n, 0.0), prev(x);
    double f = 0.0, b1t = 1.0, b2t = 1.0;
    auto eval = [&](const double* p, double* fo, double* go) {  // Adam always wants both
        if (!value(p, fo)) return false;
        ++r.grads;
        return (bool)grad(go);
    };
    for (int t = 1; t <= s.max_iters; ++t) {
        ++r.evals;
        if (!eval(x.data(), &f, g.data())) {
            if (t == 1)
                r.failed = 1;
            else
                x = prev;
            return r;
        }
        if (t == 1) r.f0 = f;
        r.f = f;
        if (below_threshold(g, s.threshold)) {
            r.converged = 1;
            return r;
        }
        prev = x;
        b1t *= s.beta1;
        b2t *= s.beta2;
        for (size_t i = 0; i < n; ++i) {
            m[i] = s.beta1 * m[i] + (1.0 - s.beta1) * g[i];
            v[i] = s.beta2 * v[i] + (1.0 - s.beta2) * g[i] * g[i];
            x[i] += s.rate * (m[i] / (1.0 - b1t)) / (std::sqrt(v[i] / (1.0 - b2t)) + s.eps);
        }
        ++r.iters;
    }
    ++r.evals;
    if (eval(x.data(), &f, g.data())) {
        if (r.iters == 0) r.f0 = f;
        r.f = f;
        r.converged = below_threshold(g, s.threshold) ? 1 : 0;
    } else if (r.iters > 0) {
        x = prev;
    } else {
        r.failed = 1;
    }
    return r;
}

// L-BFGS ascent (internally minimises -f).
template <class Value, class Grad>
OptResult lbfgs_ascent(Value&& value, Grad&& grad, std::vector<double>& x, const OptSettings& s) {
    OptResult r;
    const size_t n = x.size();
    const int hist = s.history > 0 ? s.history : 15;
    std::vector<double> g(n), xt(n), gt(n), d(n), q(n);
    double f = 0.0;
    auto neg_value = [&](const double* p, double* fo) {
        ++r.evals;
        double fv = 0.0;
        if (!value(p, &fv)) return false;
        if (!std::isfinite(fv)) return false;
        *fo = -fv;
        return true;
    };
    auto neg_grad = [&](double* go) {  // at the point neg_value was last called with
        ++r.grads;
        if (!grad(go)) return false;
        for (size_t i = 0; i < n; ++i) go[i] = -go[i];
        return true;
    };
    if (!neg_value(x.data(), &f) || !neg_grad(g.data())) {
        r.failed = 1;
        return r;
    }
    r.f0 = r.f = -f;
    if (below_threshold(g, s.threshold)) {
        r.converged = 1;
        return r;
    }
    struct Pair {
        std::vector<double> s, y;
        double rho;
    };
    std::deque<Pair> H;
    auto dot = [&](const std::vector<double>& a, const std::vector<double>& b) {
        double t = 0.0;
        for (size_t i = 0; i < n; ++i) t += a[i] * b[i];
        return t;
    };
    double gn = std::sqrt(dot(g, g));
    for (size_t i = 0; i < n; ++i) d[i] = -g[i];
    double step = gn > 0.0 ? 1.0 / gn : 1.0;
    const double c1 = 1e-4, c2 = 0.9;
    for (int it = 0; it < s.max_iters; ++it) {
        const double gd = dot(g, d);
        // bisection on the weak Wolfe conditions
        double lo = 0.0, hi = INFINITY, t = step, ft = 0.0;
        bool ok = false;
        for (int ls = 0; ls < 40; ++ls) {
            for (size_t i = 0; i < n; ++i) xt[i] = x[i] + t * d[i];
            const bool evaluated = neg_value(xt.data(), &ft);
            if (!evaluated || ft > f + c1 * t * gd || !neg_grad(gt.data())) {  // gradient only past Armijo
                hi = t;
                t = 0.5 * (lo + hi);
            } else if (dot(gt, d) < c2 * gd) {
                lo = t;
                t = std::isinf(hi) ? 2.0 * t : 0.5 * (lo + hi);
            } else {
                ok = true;
                break;
            }
            if (hi - lo < 1e-16 * (1.0 + std::fabs(hi))) break;
        }
        if (!ok) break;  // no acceptable step: keep the current point (gonum reports an error here)
        Pair p;
        p.s.resize(n);
        p.y.resize(n);
        for (size_t i = 0; i < n; ++i) {
            p.s[i] = xt[i] - x[i];
            p.y[i] = gt[i] - g[i];
        }
        const double sy = dot(p.s, p.y);
        x = xt;
        f = ft;
        g = gt;
        r.f = -f;
        ++r.iters;
        if (below_threshold(g, s.threshold)) {
            r.converged = 1;
            break;
        }
        if (sy > 1e-10 * std::sqrt(dot(p.s, p.s) * dot(p.y, p.y))) {
            p.rho = 1.0 / sy;
            H.push_back(std::move(p));
            if ((int)H.size() > hist) H.pop_front();
        }
        // two-loop recursion: d = -H g
        q = g;
        std::vector<double> al(H.size());
        for (int i = (int)H.size() - 1; i >= 0; --i) {
            al[i] = H[i].rho * dot(H[i].s, q);
            for (size_t k = 0; k < n; ++k) q[k] -= al[i] * H[i].y[k];
        }
        if (!H.empty()) {
            const Pair& last = H.back();
            const double gamma = 1.0 / (last.rho * dot(last.y, last.y));
            for (size_t k = 0; k < n; ++k) q[k] *= gamma;
        }
        for (size_t i = 0; i < H.size(); ++i) {
            const double be = H[i].rho * dot(H[i].y, q);
            for (size_t k = 0; k < n; ++k) q[k] += (al[i] - be) * H[i].s[k];
        }
        for (size_t k = 0; k < n; ++k) d[k] = -q[k];
        step = 1.0;
        if (!(dot(g, d) < 0.0)) {  // not a descent direction: restart from steepest descent
            H.clear();
            gn = std::sqrt(dot(g, g));
            for (size_t k = 0; k < n; ++k) d[k] = -g[k];
            step = gn > 0.0 ? 1.0 / gn : 1.0;
        }
    }
    return r;
}

}  // namespace gogp
