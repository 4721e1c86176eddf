// Device-side kernel-expression descriptor and its host-side lowering.
//
// The C-ABI takes a postfix program (gogp_op, include/gogp_b200.h).  The device
// evaluates a canonical form: a sum of product terms,
//     k(xa, xb) = sum_t coef_t * prod_{f in t} factor_f(xa[dim_f], xb[dim_f]; theta)
// obtained by distributing products over sums on the host.  A factor is a stock
// 1-D kernel of kernel/kernel.go (reference) or a bare parameter.  Every factor
// of the reference library is strictly positive, so the partials the reference
// gets from its AD tape (model.Gradient, gp/gp.go:113) follow from the product
// value and per-factor LOG-derivatives: d prod / d theta = prod * (d f / d theta) / f.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/gogp_b200.h"

namespace gogp {

constexpr int kMaxTerms = 8;
constexpr int kMaxFactors = 48;  // over all terms
constexpr int kMaxTheta = 32;

enum FactorKind : int {
    F_PARAM = 0,
    F_NORMAL = 1,
    F_PERIODIC = 2,
    F_MATERN32 = 3,
    F_MATERN52 = 4,  // c = 1 (shipped) or 5/3 (textbook), in `c`
    F_EVENTS = 5,    // parameter-free discount of the first event boundary between the two points
};
constexpr int kMaxEvents = 16;

struct DevFactor {
    int kind;
    int dim;
    int p0, p1;      // theta indices (p1 < 0 when unused)
    double a0, a1;   // effective parameters scale*theta, refreshed per evaluation
    double i0;       // 1 / a0
    double i1;       // pi / a1 (Periodic: the phase is |r| * i1)
    double c;        // Matern52 d^2 coefficient
};

struct DevProgram {
    int nterms;
    int ntheta;
    int fbeg[kMaxTerms + 1];
    int nnorm[kMaxTerms];  // leading Normal factors of each term (they share one exp)
    double coef[kMaxTerms];
    DevFactor f[kMaxFactors];
    int nevents;
    double ev[kMaxEvents][3];  // from, to, discount
};

struct HostFactor {
    int kind, dim, p0, p1;
    double s0, s1, c;
};
struct HostTerm {
    double coef;
    std::vector<HostFactor> f;
};

struct Program {
    int ntheta = 0;
    std::vector<HostTerm> terms;
    bool has_leaf = false;
    std::vector<double> events;  // n x 3

    // Postfix -> sum of products.  Returns false and sets err when malformed or
    // too large.
    bool lower(const gogp_op* ops, int n, int ntheta_, int ndim, bool allow_leaves, std::string* err);
    // Fill the device form for natural-scale parameters theta.
    void bind(const double* theta, DevProgram* out) const;
    // Input-independent programs (noise): value and d/d log theta_q.
    double eval_scalar(const double* theta, double* dlog /* ntheta */) const;
};

}  // namespace gogp
