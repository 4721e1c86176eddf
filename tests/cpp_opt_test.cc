// CPU unit test of the optimiser loops (gogp_b200/csrc/optimize.hpp) on analytic objectives:
// the loops are host code templated on the objective, so they are checked here without a GPU.
#include <cmath>
#include <cstdio>
#include <vector>

#include "../gogp_b200/csrc/optimize.hpp"

using namespace gogp;

static int fails = 0;
#define CHECK(c)                                                    \
    do {                                                            \
        if (!(c)) {                                                 \
            std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); \
            ++fails;                                                \
        }                                                           \
    } while (0)

int main() {
    // concave quadratic with a known maximiser: f = -0.5 sum w_i (x_i - c_i)^2
    const std::vector<double> w{1.0, 10.0, 0.1, 3.0}, c{1.0, -2.0, 0.5, 4.0};
    int calls = 0, gcalls = 0;
    std::vector<double> last(4);
    auto quad = [&](const double* x, double* f) {
        ++calls;
        double s = 0.0;
        for (size_t i = 0; i < w.size(); ++i) {
            s -= 0.5 * w[i] * (x[i] - c[i]) * (x[i] - c[i]);
            last[i] = x[i];
        }
        *f = s;
        return true;
    };
    auto quad_g = [&](double* g) {  // gradient at the point quad() was last called with
        ++gcalls;
        for (size_t i = 0; i < w.size(); ++i) g[i] = -w[i] * (last[i] - c[i]);
        return true;
    };
    {
        OptSettings s;
        s.method = 1;
        s.max_iters = 100;
        s.threshold = 1e-9;
        std::vector<double> x(4, 0.0);
        OptResult r = lbfgs_ascent(quad, quad_g, x, s);
        CHECK(r.converged == 1);
        CHECK(r.evals == calls && r.grads == gcalls && r.grads <= r.evals);
        CHECK(r.iters <= 20);
        for (size_t i = 0; i < 4; ++i) CHECK(std::fabs(x[i] - c[i]) < 1e-8);
        CHECK(std::fabs(r.f) < 1e-15 && r.f0 < -20.0);
    }
    {   // Adam: first step moves every coordinate by rate * sign(g) (bias-corrected moments)
        OptSettings s;
        s.method = 0;
        s.max_iters = 1;
        s.rate = 0.01;
        s.threshold = 1e-12;
        std::vector<double> x(4, 0.0);
        OptResult r = adam_ascent(quad, quad_g, x, s);
        CHECK(r.iters == 1 && r.evals == 2 && r.grads == 2 && r.converged == 0);
        for (size_t i = 0; i < 4; ++i) CHECK(std::fabs(std::fabs(x[i]) - 0.01) < 1e-8 && x[i] * c[i] > 0);
        s.max_iters = 20000;
        s.rate = 0.01;
        s.threshold = 1e-6;
        std::vector<double> y(4, 0.0);
        r = adam_ascent(quad, quad_g, y, s);
        CHECK(r.converged == 1);
        for (size_t i = 0; i < 4; ++i) CHECK(std::fabs(y[i] - c[i]) < 1e-4);
    }
    {   // Rosenbrock (maximise its negative): curved valley, exercises the line search
        double px = 0.0, py = 0.0;
        auto rosen = [&](const double* x, double* f) {
            const double a = 1.0 - x[0], b = x[1] - x[0] * x[0];
            *f = -(a * a + 100.0 * b * b);
            px = x[0];
            py = x[1];
            return true;
        };
        auto rosen_g = [&](double* g) {
            const double a = 1.0 - px, b = py - px * px;
            g[0] = -(-2.0 * a - 400.0 * px * b);
            g[1] = -(200.0 * b);
            return true;
        };
        OptSettings s;
        s.method = 1;
        s.max_iters = 200;
        s.threshold = 1e-8;
        std::vector<double> x{-1.2, 1.0};
        OptResult r = lbfgs_ascent(rosen, rosen_g, x, s);
        CHECK(r.converged == 1);
        CHECK(std::fabs(x[0] - 1.0) < 1e-6 && std::fabs(x[1] - 1.0) < 1e-6);
        CHECK(r.evals < 150);
        CHECK(r.grads < r.evals);  // the curved valley rejects some trial steps on their value alone
    }
    {   // a region where the objective cannot be evaluated (not positive definite): the line
        // search backs off instead of failing; a bad start is reported
        double wx = 0.0;
        auto wall = [&](const double* x, double* f) {
            if (x[0] > 2.0) return false;
            *f = -(x[0] - 1.9) * (x[0] - 1.9);
            wx = x[0];
            return true;
        };
        auto wall_g = [&](double* g) {
            g[0] = -2.0 * (wx - 1.9);
            return true;
        };
        OptSettings s;
        s.method = 1;
        s.max_iters = 50;
        s.threshold = 1e-9;
        std::vector<double> x{-50.0};
        OptResult r = lbfgs_ascent(wall, wall_g, x, s);
        CHECK(r.converged == 1 && std::fabs(x[0] - 1.9) < 1e-8);
        std::vector<double> bad{3.0};
        r = lbfgs_ascent(wall, wall_g, bad, s);
        CHECK(r.failed == 1 && bad[0] == 3.0);
        s.method = 0;
        r = adam_ascent(wall, wall_g, bad, s);
        CHECK(r.failed == 1);
        // Adam walking into the wall is rolled back to the last good point
        auto cliff = [&](const double* x, double* f) {
            if (x[0] > 0.025) return false;
            *f = x[0];
            return true;
        };
        auto cliff_g = [&](double* g) {
            g[0] = 1.0;
            return true;
        };
        s.max_iters = 10;
        s.rate = 0.01;
        std::vector<double> z{0.0};
        r = adam_ascent(cliff, cliff_g, z, s);
        CHECK(r.failed == 0 && z[0] <= 0.025 && z[0] > 0.015);
    }
    if (fails == 0) std::printf("all ok\n");
    return fails ? 1 : 0;
}
