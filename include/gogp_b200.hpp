// C++ host-side mirror of GoGP's gp.GP / gp.Model and package kernel over the
// C-ABI (gogp_b200.h).  Header only.  The reference is Go (gp/gp.go, gp/model.go,
// kernel/kernel.go, kernel/noise.go); where no Go toolchain exists this is the
// compiled-language host layer: same names, same argument layout
// ([log theta | X flat | Y]), same error split (Observe throws where the
// reference panics, Absorb / Produce report an error).
//
//   using namespace gogp;
//   GP g(1, Param(0) * Matern32.Of(1), 0.01 * UniformNoise());   // tutorial/barebones
//   g.X = ...; g.Y = ...;
//   double ll = g.Observe(x);  std::vector<double> dll = g.Gradient();
#pragma once
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "gogp_b200.h"

namespace gogp {

// ---- package kernel --------------------------------------------------------------
// A kernel is an expression over the stock kernels; it carries its postfix
// descriptor (a user Simil written in host code cannot run on the device).
class Kernel {
public:
    std::vector<gogp_op> ops;
    int ntheta = 0;

    int NTheta() const { return ntheta; }  // gp.Kernel.NTheta, gp/gp.go:16
    Kernel WithNTheta(int n) const {       // tutorial/anynoise/kernel/kernel.go:31-35
        Kernel k = *this;
        k.ntheta = n;
        return k;
    }
    static Kernel Binary(const Kernel& a, const Kernel& b, gogp_op_kind kind) {
        Kernel k;
        k.ops = a.ops;
        k.ops.insert(k.ops.end(), b.ops.begin(), b.ops.end());
        gogp_op o{};
        o.kind = (uint8_t)kind;
        k.ops.push_back(o);
        k.ntheta = a.ntheta > b.ntheta ? a.ntheta : b.ntheta;
        return k;
    }
};

inline Kernel Const(double c) {
    Kernel k;
    gogp_op o{};
    o.kind = GOGP_OP_CONST;
    o.constant = c;
    k.ops.push_back(o);
    return k;
}
inline Kernel Param(int i, double scale = 1.0) {
    Kernel k;
    gogp_op o{};
    o.kind = GOGP_OP_PARAM;
    o.param[0] = (int16_t)i;
    o.scale[0] = scale;
    o.scale[1] = 1.0;
    k.ops.push_back(o);
    k.ntheta = i + 1;
    return k;
}
inline Kernel operator+(const Kernel& a, const Kernel& b) { return Kernel::Binary(a, b, GOGP_OP_ADD); }
inline Kernel operator*(const Kernel& a, const Kernel& b) { return Kernel::Binary(a, b, GOGP_OP_MUL); }
inline Kernel operator*(double c, const Kernel& b) { return Const(c) * b; }

// Stock 1-D similarity kernels (kernel/kernel.go:13-92).  As a value they have the
// reference's parameter layout ([l] or [l, p]); Of() re-targets the parameter slots
// (with constant multipliers) and the input coordinate.
class Stock : public Kernel {
public:
    gogp_op_kind kind;
    explicit Stock(gogp_op_kind k) : kind(k) { *static_cast<Kernel*>(this) = Of(0, 1, 0); }
    Kernel Of(int l, int p = 0, int dim = 0, double lscale = 1.0, double pscale = 1.0) const {
        Kernel k;
        gogp_op o{};
        o.kind = (uint8_t)kind;
        o.dim = (uint8_t)dim;
        o.param[0] = (int16_t)l;
        o.param[1] = (int16_t)(kind == GOGP_OP_PERIODIC ? p : 0);
        o.scale[0] = lscale;
        o.scale[1] = pscale;
        k.ops.push_back(o);
        k.ntheta = (kind == GOGP_OP_PERIODIC && p > l ? p : l) + 1;
        return k;
    }
};
static const Stock Normal(GOGP_OP_NORMAL);      // kernel/kernel.go:13-26
static const Stock Periodic(GOGP_OP_PERIODIC);  // kernel/kernel.go:34-47
static const Stock Matern32(GOGP_OP_MATERN32);  // kernel/kernel.go:60-73
static const Stock Matern52(GOGP_OP_MATERN52);  // kernel/kernel.go:79-92 (5/3 == 1 as shipped)
inline Kernel ConstantNoise(double std) { return Const(std * std); }  // kernel/noise.go:21-34
inline Kernel UniformNoise() { return Param(0) * Param(0); }          // kernel/noise.go:39-53

// ---- package gp ---------------------------------------------------------------------
struct Error {  // an `error` value of the reference; ok() == (err == nil)
    gogp_status status = GOGP_OK;
    std::string message;
    bool ok() const { return status == GOGP_OK; }
};

class Panic : public std::runtime_error {  // where the reference panics (gp/gp.go:398-405)
public:
    gogp_status status;
    Panic(gogp_status s, const std::string& m) : std::runtime_error(m), status(s) {}
};

class GP {  // gp.GP, gp/gp.go:20-38
public:
    int NDim;
    Kernel Simil, Noise;
    bool HasNoise;
    std::vector<double> ThetaSimil, ThetaNoise;
    std::vector<std::vector<double>> X;
    std::vector<double> Y;
    bool Parallel = false;  // kept for API parity (gp/gp.go:31); the GPU build is the parallel path
    int Device = 0;

    GP(int ndim, Kernel simil) : NDim(ndim), Simil(std::move(simil)), HasNoise(false) {}
    GP(int ndim, Kernel simil, Kernel noise)
        : NDim(ndim), Simil(std::move(simil)), Noise(std::move(noise)), HasNoise(true) {}
    GP(const GP&) = delete;
    GP& operator=(const GP&) = delete;
    ~GP() {
        if (h_) gogp_destroy(h_);
    }

    // gp/gp.go:80-87
    Error Absorb(const std::vector<std::vector<double>>& x, const std::vector<double>& y) {
        defaults();
        X = x;
        Y = y;
        Error e = handle();
        if (!e.ok()) return e;
        std::vector<double> xf = flatten(x);
        // the C-ABI takes one N for both buffers: a ragged x or len(x) != len(y) must not reach it
        if (xf.size() != y.size() * (size_t)NDim) return bad("len(x) != len(y) (or a row of x is not NDim long)");
        withObs_ = false;
        n_ = (int64_t)y.size();
        return wrap(gogp_absorb(h_, ThetaSimil.data(), ThetaNoise.data(), xf.data(), y.data(), n_));
    }

    // the growing window of tutorial.Evaluate (tutorial/tutorial.go:91-179) at unchanged hyper-parameters: the factor
    // is extended instead of recomputed; equivalent to Absorb on the concatenated data
    Error Extend(const std::vector<std::vector<double>>& x, const std::vector<double>& y) {
        Error e = handle();
        if (!e.ok()) return e;
        std::vector<double> xf = flatten(x);
        if (xf.size() != y.size() * (size_t)NDim) return bad("len(x) != len(y) (or a row of x is not NDim long)");
        double lml = 0.0;
        e = wrap(gogp_extend(h_, xf.data(), y.data(), (int64_t)y.size(), &lml));
        if (!e.ok()) return e;
        X.insert(X.end(), x.begin(), x.end());
        Y.insert(Y.end(), y.begin(), y.end());
        n_ += (int64_t)y.size();
        return e;
    }

    // gp/gp.go:35-36, 255-257: the exported results a user may store, and "Produce works on stored results"
    std::vector<double> Alpha() {
        std::vector<double> a((size_t)n_, 0.0);
        if (h_ && n_ > 0 && gogp_get_alpha(h_, a.data(), n_) != GOGP_OK) throw Panic(GOGP_NOT_READY, gogp_last_error(h_));
        return a;
    }
    std::vector<double> L() {  // n x n row-major, lower
        std::vector<double> l((size_t)n_ * (size_t)n_, 0.0);
        if (h_ && n_ > 0 && gogp_get_factor(h_, l.data(), n_) != GOGP_OK) throw Panic(GOGP_NOT_READY, gogp_last_error(h_));
        return l;
    }
    Error Restore(const std::vector<double>& thetaSimil, const std::vector<double>& thetaNoise,
                  const std::vector<std::vector<double>>& x, const std::vector<double>& alpha,
                  const std::vector<double>& l) {
        Error e = handle();
        if (!e.ok()) return e;
        std::vector<double> xf = flatten(x);
        const size_t n = alpha.size();
        if (xf.size() != n * (size_t)NDim || l.size() != n * n || thetaSimil.size() != (size_t)Simil.NTheta() ||
            thetaNoise.size() != (size_t)noiseNTheta())
            return bad("state shapes do not agree");
        ThetaSimil = thetaSimil;
        ThetaNoise = thetaNoise;
        X = x;
        withObs_ = false;
        n_ = (int64_t)n;
        return wrap(gogp_set_state(h_, ThetaSimil.data(), ThetaNoise.data(), xf.data(), n_, alpha.data(), l.data()));
    }

    // gp/gp.go:244-253
    double LML() {
        double out = 0.0;
        if (!h_ || gogp_lml(h_, &out) != GOGP_OK) return 0.0;
        return out;
    }

    // gp/gp.go:258-360
    Error Produce(const std::vector<std::vector<double>>& x, std::vector<double>& mu, std::vector<double>& sigma) {
        defaults();
        Error e = handle();
        if (!e.ok()) return e;
        std::vector<double> zf = flatten(x);
        if (zf.size() != x.size() * (size_t)NDim) return bad("a row of x is not NDim long");
        mu.assign(x.size(), 0.0);
        sigma.assign(x.size(), 0.0);
        return wrap(gogp_produce(h_, zf.data(), (int64_t)x.size(), mu.data(), sigma.data()));
    }

    // gp/gp.go:374-413.  x = [log theta] or [log theta | X flat | Y]; the parameter prefix is
    // exponentiated in place during the call and logged back, as in the reference.
    double Observe(std::vector<double>& x) {
        defaults();
        Error e = handle();
        if (!e.ok()) throw Panic(e.status, e.message);
        const size_t P = (size_t)Simil.NTheta() + (size_t)noiseNTheta();
        if (x.size() < P) throw Panic(GOGP_BAD_ARGUMENT, "len(x)");
        std::vector<double> logTheta(x.begin(), x.begin() + (long)P);
        for (size_t i = 0; i < P; ++i) x[i] = std::exp(x[i]);
        struct Restore {
            std::vector<double>& x;
            size_t P;
            ~Restore() {
                for (size_t i = 0; i < P; ++i) x[i] = std::log(x[i]);
            }
        } restore{x, P};
        ThetaSimil.assign(x.begin(), x.begin() + Simil.NTheta());
        ThetaNoise.assign(x.begin() + Simil.NTheta(), x.begin() + (long)P);
        const size_t rest = x.size() - P;
        withObs_ = rest > 0;
        std::vector<double> xf;
        const double *xp, *yp;
        if (withObs_) {
            const size_t n = rest / (size_t)(NDim + 1);
            if (n * (size_t)(NDim + 1) != rest) throw Panic(GOGP_BAD_ARGUMENT, "len(x)");  // gp/gp.go:398-400
            n_ = (int64_t)n;
            xp = x.data() + P;
            yp = xp + n * (size_t)NDim;
            X.assign(n, std::vector<double>());
            for (size_t i = 0; i < n; ++i) X[i].assign(xp + i * (size_t)NDim, xp + (i + 1) * (size_t)NDim);
            Y.assign(yp, yp + n);
        } else {
            xf = flatten(X);
            if (xf.size() != Y.size() * (size_t)NDim) throw Panic(GOGP_BAD_ARGUMENT, "len(gp.X) != len(gp.Y)");
            n_ = (int64_t)Y.size();
            xp = xf.data();
            yp = Y.data();
        }
        double lml = 0.0;
        gogp_status st = gogp_observe(h_, logTheta.data(), withObs_ ? 1 : 0, xp, yp, n_, &lml);
        if (st != GOGP_OK) throw Panic(st, gogp_last_error(h_));  // panic(err), gp/gp.go:403-405
        return lml;
    }

    // gp/gp.go:418-499
    std::vector<double> Gradient() {
        size_t n = (size_t)Simil.NTheta() + (size_t)noiseNTheta();
        if (withObs_) n += (size_t)n_ * (size_t)(NDim + 1);
        std::vector<double> grad(n, 0.0);
        if (!h_ || n == 0) return grad;
        gogp_status st = gogp_gradient(h_, grad.data(), (int64_t)n);
        if (st != GOGP_OK) throw Panic(st, gogp_last_error(h_));
        return grad;
    }

    // tutorial/tutorial.go:124-175 inside the library (gogp_optimize): alg 0 = the infer.Adam loop,
    // 1 = L-BFGS; x (log hyper-parameters) is updated in place over gp.X, gp.Y.  prior may be null.
    gogp_opt_result Optimize(std::vector<double>& x, int alg, int iters, double threshold, double rate = 0.01,
                             gogp_prior_fn prior = nullptr, void* ctx = nullptr) {
        defaults();
        Error e = handle();
        if (!e.ok()) throw Panic(e.status, e.message);
        const size_t P = (size_t)Simil.NTheta() + (size_t)noiseNTheta();
        if (x.size() != P) throw Panic(GOGP_BAD_ARGUMENT, "len(x)");
        std::vector<double> xf = flatten(X);
        if (xf.size() != Y.size() * (size_t)NDim) throw Panic(GOGP_BAD_ARGUMENT, "len(gp.X) != len(gp.Y)");
        gogp_status st = gogp_set_data(h_, xf.data(), Y.data(), (int64_t)Y.size());
        if (st != GOGP_OK) throw Panic(st, gogp_last_error(h_));
        gogp_opt_settings s{};
        s.method = alg;
        s.max_iters = iters;
        s.threshold = threshold;
        s.rate = rate;
        gogp_opt_result r{};
        st = gogp_optimize(h_, &s, x.data(), prior, ctx, &r);
        if (st != GOGP_OK) throw Panic(st, gogp_last_error(h_));
        withObs_ = false;
        n_ = (int64_t)Y.size();
        ThetaSimil.resize((size_t)Simil.NTheta());
        ThetaNoise.resize((size_t)noiseNTheta());
        for (size_t i = 0; i < P; ++i)
            (i < ThetaSimil.size() ? ThetaSimil[i] : ThetaNoise[i - ThetaSimil.size()]) = std::exp(x[i]);
        return r;
    }

private:
    gogp_handle* h_ = nullptr;
    bool withObs_ = false;
    int64_t n_ = 0;

    int noiseNTheta() const { return HasNoise ? Noise.NTheta() : 0; }
    void defaults() {  // gp/gp.go:45-57
        if (ThetaSimil.empty()) ThetaSimil.assign((size_t)Simil.NTheta(), 0.0);
        if (ThetaNoise.empty()) ThetaNoise.assign((size_t)noiseNTheta(), 0.0);
    }
    std::vector<double> flatten(const std::vector<std::vector<double>>& x) const {
        std::vector<double> out;
        out.reserve(x.size() * (size_t)NDim);
        for (const auto& r : x) out.insert(out.end(), r.begin(), r.end());
        return out;
    }
    static Error bad(const char* msg) {
        Error e;
        e.status = GOGP_BAD_ARGUMENT;
        e.message = msg;
        return e;
    }
    Error wrap(gogp_status st) const {
        Error e;
        e.status = st;
        if (st != GOGP_OK) e.message = gogp_last_error(h_);
        return e;
    }
    Error handle() {
        if (h_) return Error();
        gogp_status st = gogp_create(NDim, Simil.ops.data(), (int)Simil.ops.size(), Simil.NTheta(),
                                     HasNoise ? Noise.ops.data() : nullptr, HasNoise ? (int)Noise.ops.size() : 0,
                                     noiseNTheta(), Device, &h_);
        Error e;
        e.status = st;
        if (st != GOGP_OK) {
            e.message = h_ ? gogp_last_error(h_) : "gogp_create failed";
            if (h_) gogp_destroy(h_);
            h_ = nullptr;
        }
        return e;
    }
};

// gp.Model (gp/model.go:9-28): GP plus priors; Priors is any type with
// double Observe(std::vector<double>&) and std::vector<double> Gradient().
template <class Priors>
class Model {
public:
    GP* gp;
    Priors* priors;
    Model(GP* g, Priors* p) : gp(g), priors(p) {}
    double Observe(std::vector<double>& x) {
        const double gll = gp->Observe(x);
        gGrad_ = gp->Gradient();
        const double pll = priors->Observe(x);
        pGrad_ = priors->Gradient();
        return gll + pll;
    }
    std::vector<double> Gradient() {
        for (size_t i = 0; i < pGrad_.size(); ++i) gGrad_[i] += pGrad_[i];
        return gGrad_;
    }

private:
    std::vector<double> gGrad_, pGrad_;
};

}  // namespace gogp
