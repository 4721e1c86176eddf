// C++ host mirror (include/gogp_b200.hpp) against the reference's TestElementalModel and
// TestProduce tables (gp/gp_test.go:23-120, 180-229).  "nodevice" mode: expect the loud
// failure the product path must give without a GPU.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "../include/gogp_b200.hpp"

using namespace gogp;

static int fails = 0;
#define CHECK(c)                                                 \
    do {                                                         \
        if (!(c)) {                                              \
            std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); \
            ++fails;                                             \
        }                                                        \
    } while (0)

int main(int argc, char** argv) {
    if (argc > 1 && !std::strcmp(argv[1], "nodevice")) {
        GP g(1, Normal, ConstantNoise(0.1));
        std::vector<double> x{0.0};
        try {
            g.Observe(x);
        } catch (const Panic& p) {
            std::printf("panic: %s\n", p.what());
            return p.status == GOGP_CUDA_ERROR ? 0 : 2;
        }
        return 1;
    }
    struct E { const char* name; bool uniform; double noise; std::vector<double> x; double ll; };
    std::vector<E> cases = {
        {"prior", false, 0.0, {0}, 0.0},
        {"single", false, 0.0, {0, 0, 1}, -1.418939},
        {"nonoise", false, 0.0, {0, 0, 1, 1, 0}, -2.399528},
        {"withnoise", false, 0.1, {1, -2, -1, 1, 0}, -4.321055},
        {"uninoise", true, 0.0, {1, 1, -1, -1, 1, 0}, -4.018110},
    };
    for (auto& c : cases) {
        GP g(1, Normal, c.uniform ? UniformNoise() : ConstantNoise(c.noise));
        std::vector<double> x = c.x;
        const double ll = g.Observe(x);
        std::vector<double> dll = g.Gradient();
        CHECK(std::fabs(ll - c.ll) < 1e-6);
        CHECK(dll.size() == c.x.size());
        for (size_t j = 0; j < x.size(); ++j) {  // forward difference, gp/gp_test.go:242-252
            std::vector<double> xj = c.x;
            xj[j] += 1e-8;
            const double llj = g.Observe(xj);
            CHECK(std::fabs(dll[j] - (llj - ll) / 1e-8) <= 1e-4);
        }
        for (size_t j = 0; j < x.size(); ++j) CHECK(std::fabs(x[j] - c.x[j]) < 1e-15);  // exp/log round trip
        std::printf("%s ll=%.6f ok\n", c.name, ll);
    }
    {  // TestProduce "noise", gp/gp_test.go:108-120
        GP g(1, Normal, ConstantNoise(0.1));
        g.ThetaSimil = {1.0};
        Error e = g.Absorb({{0}, {1}}, {1, -1});
        CHECK(e.ok());
        std::vector<double> mu, sigma;
        e = g.Produce({{-2.}, {3.}}, mu, sigma);
        CHECK(e.ok());
        CHECK(std::fabs(mu[0] - 0.307895) < 1e-6 && std::fabs(mu[1] + 0.307895) < 1e-6);
        CHECK(std::fabs(sigma[0] - 0.987037) < 1e-6 && std::fabs(sigma[1] - 0.987037) < 1e-6);
        GP bad(1, Normal, ConstantNoise(0.0));
        bad.ThetaSimil = {1.0};
        e = bad.Absorb({{0}, {0}}, {1, 2});  // singular K: an error, not a panic, from Absorb
        CHECK(!e.ok() && e.status == GOGP_NOT_POSITIVE_DEFINITE);
        // len(x) != len(y) and a ragged x are errors of the wrapper: the C-ABI takes one N for both buffers
        e = bad.Absorb({{0}, {1}, {2}}, {1, 2});
        CHECK(!e.ok() && e.status == GOGP_BAD_ARGUMENT);
        e = bad.Absorb({{0}, {1, 5}}, {1, 2});
        CHECK(!e.ok() && e.status == GOGP_BAD_ARGUMENT);
        // "Produce works on stored results" (gp/gp.go:255-257): restore L and Alpha into a fresh GP
        std::vector<double> alpha = g.Alpha(), l = g.L();
        CHECK(alpha.size() == 2 && l.size() == 4 && l[1] == 0.0);
        GP r(1, Normal, ConstantNoise(0.1));
        e = r.Restore({1.0}, {}, {{0}, {1}}, alpha, l);
        CHECK(e.ok());
        std::vector<double> mu2, sigma2;
        e = r.Produce({{-2.}, {3.}}, mu2, sigma2);
        CHECK(e.ok());
        CHECK(std::fabs(mu2[0] - mu[0]) < 1e-12 && std::fabs(sigma2[1] - sigma[1]) < 1e-12);
        e = r.Restore({1.0}, {}, {{0}, {1}}, alpha, {1.0});
        CHECK(!e.ok() && e.status == GOGP_BAD_ARGUMENT);
        // the growing window: Absorb one point, Extend by the other == Absorb both
        GP w(1, Normal, ConstantNoise(0.1));
        w.ThetaSimil = {1.0};
        e = w.Absorb({{0}}, {1});
        CHECK(e.ok());
        e = w.Extend({{1}}, {-1});
        CHECK(e.ok());
        CHECK(std::fabs(w.LML() - g.LML()) < 1e-12);
        std::vector<double> mu3, sigma3;
        e = w.Produce({{-2.}, {3.}}, mu3, sigma3);
        CHECK(e.ok() && std::fabs(mu3[0] - mu[0]) < 1e-12 && std::fabs(sigma3[0] - sigma[0]) < 1e-12);
    }
    std::printf(fails ? "FAILED %d\n" : "all ok\n", fails);
    return fails ? 1 : 0;
}
