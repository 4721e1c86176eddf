// TEST INFRASTRUCTURE: evaluates a kernel descriptor on the HOST with the product's own lowering
// (gogp_b200/csrc/program.cc) and element functions (gogp_b200/csrc/kexpr.cuh, compiled here as plain
// inline C++), combined exactly as the device kernels combine them: value as in cov_tile_kernel,
// log-parameter partials as in grad_trace_kernel's phase 2, input partials as in grad_inputs_kernel.
// Lets the CPU test tier check descriptor lowering and the closed-form partials against the oracle
// without a GPU.  Never linked into libgogp_b200.so.
#include <string>

#include "../gogp_b200/csrc/kexpr.cuh"
#include "../gogp_b200/csrc/program.cc"

using namespace gogp;

extern "C" int cpu_kexpr_eval(const gogp_op* ops, int nops, int ntheta, int ndim, const double* theta,
                              const double* events, int nev, const double* xa, const double* xb, double* value,
                              double* dlog, double* dxa) {
    Program p;
    std::string err;
    if (!p.lower(ops, nops, ntheta, ndim, true, &err)) return 1;
    if (nev > 0) p.events.assign(events, events + 3 * nev);
    DevProgram d;
    p.bind(theta, &d);
    auto XA = [&](int k) { return xa[k]; };
    auto XB = [&](int k) { return xb[k]; };
    double v = 0.0;
    for (int q = 0; q < ntheta; ++q) dlog[q] = 0.0;
    for (int k = 0; k < ndim; ++k) dxa[k] = 0.0;
    for (int t = 0; t < d.nterms; ++t) {
        const double P = term_value(d, t, XA, XB);
        v += P;
        for (int fi = d.fbeg[t]; fi < d.fbeg[t + 1]; ++fi) {
            const DevFactor& f = d.f[fi];
            if (f.p0 >= 0) {
                double g0, g1;
                factor_dlog_theta(f, xa[f.dim], xb[f.dim], g0, g1);
                dlog[f.p0] += P * g0;
                if (f.p1 >= 0) dlog[f.p1] += P * g1;
            }
        }
        for (int k = 0; k < ndim; ++k) {
            double g = 0.0;
            for (int fi = d.fbeg[t]; fi < d.fbeg[t + 1]; ++fi) {
                const DevFactor& f = d.f[fi];
                if (f.dim == k && f.kind != F_PARAM) g += factor_dlog_xa(f, xa[k], xb[k]);
            }
            if (g != 0.0) dxa[k] += P * g;
        }
    }
    *value = v;
    return 0;
}
