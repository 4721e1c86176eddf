// Host-side lowering of the postfix kernel descriptor (see program.h).
#include "program.h"

#include <cstdio>

namespace gogp {

bool Program::lower(const gogp_op* ops, int n, int ntheta_, int ndim, bool allow_leaves, std::string* err) {
    ntheta = ntheta_;
    terms.clear();
    has_leaf = false;
    if (ntheta < 0 || ntheta > kMaxTheta) {
        *err = "ntheta out of range";
        return false;
    }
    std::vector<std::vector<HostTerm>> stack;
    auto bad = [&](const char* msg, int i) {
        char buf[128];
        snprintf(buf, sizeof buf, "kernel descriptor: %s at op %d", msg, i);
        *err = buf;
        return false;
    };
    for (int i = 0; i < n; ++i) {
        const gogp_op& op = ops[i];
        switch (op.kind) {
            case GOGP_OP_CONST:
                stack.push_back({HostTerm{op.constant, {}}});
                break;
            case GOGP_OP_PARAM: {
                if (op.param[0] < 0 || op.param[0] >= ntheta) return bad("parameter index out of range", i);
                HostTerm t{op.scale[0], {}};
                t.f.push_back(HostFactor{F_PARAM, 0, op.param[0], -1, 1.0, 1.0, 0.0});
                stack.push_back({t});
                break;
            }
            case GOGP_OP_NORMAL:
            case GOGP_OP_PERIODIC:
            case GOGP_OP_MATERN32:
            case GOGP_OP_MATERN52:
            case GOGP_OP_MATERN52_TEXTBOOK: {
                if (!allow_leaves) return bad("similarity leaf in a noise program", i);
                if (op.dim >= ndim) return bad("input dimension out of range", i);
                if (op.param[0] < 0 || op.param[0] >= ntheta) return bad("parameter index out of range", i);
                HostFactor f{};
                f.dim = op.dim;
                f.p0 = op.param[0];
                f.p1 = -1;
                f.s0 = op.scale[0];
                f.s1 = 1.0;
                f.c = 0.0;
                if (op.kind == GOGP_OP_NORMAL) f.kind = F_NORMAL;
                if (op.kind == GOGP_OP_MATERN32) f.kind = F_MATERN32;
                if (op.kind == GOGP_OP_MATERN52) { f.kind = F_MATERN52; f.c = 1.0; }
                if (op.kind == GOGP_OP_MATERN52_TEXTBOOK) { f.kind = F_MATERN52; f.c = 5.0 / 3.0; }
                if (op.kind == GOGP_OP_PERIODIC) {
                    f.kind = F_PERIODIC;
                    if (op.param[1] < 0 || op.param[1] >= ntheta) return bad("period index out of range", i);
                    f.p1 = op.param[1];
                    f.s1 = op.scale[1];
                }
                HostTerm t{1.0, {f}};
                stack.push_back({t});
                has_leaf = true;
                break;
            }
            case GOGP_OP_EVENTS: {
                if (!allow_leaves) return bad("similarity leaf in a noise program", i);
                if (op.dim >= ndim) return bad("input dimension out of range", i);
                HostTerm t{1.0, {HostFactor{F_EVENTS, op.dim, -1, -1, 1.0, 1.0, 0.0}}};
                stack.push_back({t});
                has_leaf = true;
                break;
            }
            case GOGP_OP_ADD:
            case GOGP_OP_MUL: {
                if (stack.size() < 2) return bad("stack underflow", i);
                std::vector<HostTerm> b = std::move(stack.back());
                stack.pop_back();
                std::vector<HostTerm> a = std::move(stack.back());
                stack.pop_back();
                std::vector<HostTerm> r;
                if (op.kind == GOGP_OP_ADD) {
                    r = std::move(a);
                    r.insert(r.end(), b.begin(), b.end());
                } else {
                    for (const HostTerm& ta : a)
                        for (const HostTerm& tb : b) {
                            HostTerm t{ta.coef * tb.coef, ta.f};
                            t.f.insert(t.f.end(), tb.f.begin(), tb.f.end());
                            r.push_back(std::move(t));
                        }
                }
                if ((int)r.size() > kMaxTerms) {
                    *err = "kernel descriptor expands to too many product terms";
                    return false;
                }
                stack.push_back(std::move(r));
                break;
            }
            default:
                return bad("unknown op kind", i);
        }
    }
    if (stack.size() != 1) {
        *err = "kernel descriptor: program must leave exactly one value";
        return false;
    }
    terms = std::move(stack.back());
    // Normal factors first within each term (stable): the device folds them into one exponential
    for (HostTerm& t : terms) {
        std::vector<HostFactor> a, b;
        for (const HostFactor& f : t.f) (f.kind == F_NORMAL ? a : b).push_back(f);
        a.insert(a.end(), b.begin(), b.end());
        t.f = std::move(a);
    }
    size_t nf = 0;
    for (const HostTerm& t : terms) nf += t.f.size();
    if (nf > (size_t)kMaxFactors) {
        *err = "kernel descriptor expands to too many factors";
        return false;
    }
    return true;
}

void Program::bind(const double* theta, DevProgram* out) const {
    out->nterms = (int)terms.size();
    out->ntheta = ntheta;
    int k = 0;
    for (int t = 0; t < out->nterms; ++t) {
        out->fbeg[t] = k;
        out->coef[t] = terms[t].coef;
        out->nnorm[t] = 0;
        for (const HostFactor& f : terms[t].f) out->nnorm[t] += f.kind == F_NORMAL;
        for (const HostFactor& f : terms[t].f) {
            DevFactor& d = out->f[k++];
            d.kind = f.kind;
            d.dim = f.dim;
            d.p0 = f.p0;
            d.p1 = f.p1;
            d.a0 = f.p0 >= 0 ? f.s0 * theta[f.p0] : 1.0;
            d.a1 = f.p1 >= 0 ? f.s1 * theta[f.p1] : 0.0;
            d.i0 = 1.0 / d.a0;
            d.i1 = d.a1 != 0.0 ? 3.14159265358979323846 / d.a1 : 0.0;
            d.c = f.c;
        }
    }
    out->fbeg[out->nterms] = k;
    out->nevents = (int)(events.size() / 3);
    for (int e = 0; e < out->nevents; ++e)
        for (int c = 0; c < 3; ++c) out->ev[e][c] = events[3 * e + c];
}

double Program::eval_scalar(const double* theta, double* dlog) const {
    for (int q = 0; q < ntheta; ++q) dlog[q] = 0.0;
    double v = 0.0;
    for (const HostTerm& t : terms) {
        double p = t.coef;
        for (const HostFactor& f : t.f) p *= theta[f.p0];  // only F_PARAM here
        v += p;
        // d p / d log theta_q = p * (multiplicity of q in the term)
        for (const HostFactor& f : t.f) dlog[f.p0] += p;
    }
    return v;
}

}  // namespace gogp
