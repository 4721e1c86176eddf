// Device-side evaluation of the canonical kernel expression (program.h).
//
// Values follow the reference's arithmetic operation by operation
// (kernel/kernel.go:23-26, 44-47, 70-73, 89-92); partials are closed forms of
// what the reference's AD tape returns (kernel/ad/kernel.go), expressed as
// log-derivatives so a product term needs only its value:
//     d(prod)/d log(theta_q) = prod * theta_q * (d f/d theta_q) / f
#pragma once
#include <cmath>

#include "program.h"

// The element functions are plain arithmetic: under nvcc they are host + device (the kernels
// use the device side), under a host compiler they are ordinary inline functions, so the CPU
// test tier can evaluate the lowered program with exactly this code (tests/cpu_kexpr_backend.cc).
#if defined(__CUDACC__)
#define GOGP_HD __host__ __device__ __forceinline__
#else
#define GOGP_HD inline
#endif

namespace gogp {

#define GOGP_SQRT3 1.7320508075688772
#define GOGP_SQRT5 2.2360679774997900
#define GOGP_PI 3.14159265358979323846

// Length-scale and period divisions use the host-computed reciprocals i0 = 1/(scale*theta),
// i1 = pi/(scale*theta_p) (d differs from the reference's r/l by at most an ulp or two).
// tutorial/events/kernel/kernel.go:33-44: order the pair, the first event whose from- or to-boundary
// lies in (lo, hi] discounts the similarity.
GOGP_HD double events_value(const DevProgram& prog, double xa, double xb) {
    const double lo = xa > xb ? xb : xa, hi = xa > xb ? xa : xb;
    for (int e = 0; e < prog.nevents; ++e) {
        const double from = prog.ev[e][0], to = prog.ev[e][1];
        if ((lo < from && from <= hi) || (lo < to && to <= hi)) return prog.ev[e][2];
    }
    return 1.0;
}

GOGP_HD double factor_value(const DevFactor& f, double xa, double xb) {
    switch (f.kind) {
        case F_PARAM:
            return f.a0;
        case F_NORMAL: {
            double d = (xa - xb) * f.i0;
            return exp(-d * d / 2);
        }
        case F_PERIODIC: {
            double d = sin(fabs(xa - xb) * f.i1) * f.i0;
            return exp(-2 * d * d);
        }
        case F_MATERN32: {
            double d = fabs(xa - xb) * f.i0;
            return (1 + GOGP_SQRT3 * d) * exp(-GOGP_SQRT3 * d);
        }
        default: {  // F_MATERN52
            double d = fabs(xa - xb) * f.i0;
            return (1 + GOGP_SQRT5 * d + f.c * d * d) * exp(-GOGP_SQRT5 * d);
        }
    }
}

// Value of product term t at (xa, xb); XA / XB map an input dimension to the
// coordinate.  The leading Normal factors of a term (the host sorts them first)
// share one exponential: prod_d exp(-d_d^2/2) = exp(-(sum_d d_d^2)/2).
template <class XA, class XB>
GOGP_HD double term_value(const DevProgram& prog, int t, XA xa, XB xb) {
    double p = prog.coef[t];
    int fi = prog.fbeg[t];
    const int fn = fi + prog.nnorm[t], fe = prog.fbeg[t + 1];
    if (fi < fn) {
        double q = 0.0;
        for (; fi < fn; ++fi) {
            const DevFactor& f = prog.f[fi];
            const double d = (xa(f.dim) - xb(f.dim)) * f.i0;
            q = fma(d, d, q);
        }
        p *= exp(-q / 2);
    }
    for (; fi < fe; ++fi) {
        const DevFactor& f = prog.f[fi];
        p *= (f.kind == F_EVENTS) ? events_value(prog, xa(f.dim), xb(f.dim)) : factor_value(f, xa(f.dim), xb(f.dim));
    }
    return p;
}

// theta_q * (d f / d theta_q) / f for the (up to) two parameters of a factor.
GOGP_HD void factor_dlog_theta(const DevFactor& f, double xa, double xb, double& g0, double& g1) {
    g1 = 0.0;
    switch (f.kind) {
        case F_EVENTS:
            g0 = 0.0;
            return;
        case F_PARAM:
            g0 = 1.0;
            return;
        case F_NORMAL: {
            double d = (xa - xb) * f.i0;
            g0 = d * d;
            return;
        }
        case F_PERIODIC: {
            double u = fabs(xa - xb) * f.i1;
            double s, c;
            sincos(u, &s, &c);
            double d = s * f.i0;
            g0 = 4 * d * d;
            g1 = 4 * d * c * u * f.i0;
            return;
        }
        case F_MATERN32: {
            double d = fabs(xa - xb) * f.i0;
            g0 = 3 * d * d / (1 + GOGP_SQRT3 * d);
            return;
        }
        default: {
            double d = fabs(xa - xb) * f.i0;
            g0 = d * d * (5 - 2 * f.c + GOGP_SQRT5 * f.c * d) / (1 + GOGP_SQRT5 * d + f.c * d * d);
            return;
        }
    }
}

// 1 / t for t >= 1 (the Matern polynomials): the hardware reciprocal + Newton steps, without the slow path of a
// general FP64 division
GOGP_HD double rcp_pos(double t) {
#if defined(__CUDA_ARCH__)
    return __drcp_rn(t);
#else
    return 1.0 / t;
#endif
}

// Value and theta log-derivatives of one factor in one go: the gradient trace needs both, and they share their
// transcendental (one sincos + one exp for Periodic, one exp for the Materns).  Values exactly as factor_value.
GOGP_HD void factor_value_dlog(const DevProgram& prog, const DevFactor& f, double xa, double xb, double& val, double& g0,
                               double& g1) {
    g1 = 0.0;
    switch (f.kind) {
        case F_EVENTS:
            val = events_value(prog, xa, xb);
            g0 = 0.0;
            return;
        case F_PARAM:
            val = f.a0;
            g0 = 1.0;
            return;
        case F_NORMAL: {
            const double d = (xa - xb) * f.i0;
            val = exp(-d * d / 2);
            g0 = d * d;
            return;
        }
        case F_PERIODIC: {
            const double u = fabs(xa - xb) * f.i1;
            double sn, cs;
            sincos(u, &sn, &cs);
            const double d = sn * f.i0;
            val = exp(-2 * d * d);
            g0 = 4 * d * d;
            g1 = 4 * d * cs * u * f.i0;
            return;
        }
        case F_MATERN32: {
            const double d = fabs(xa - xb) * f.i0;
            const double t = 1 + GOGP_SQRT3 * d;
            val = t * exp(-GOGP_SQRT3 * d);
            g0 = 3 * d * d * rcp_pos(t);
            return;
        }
        default: {  // F_MATERN52
            const double d = fabs(xa - xb) * f.i0;
            const double t = 1 + GOGP_SQRT5 * d + f.c * d * d;
            val = t * exp(-GOGP_SQRT5 * d);
            g0 = d * d * (5 - 2 * f.c + GOGP_SQRT5 * f.c * d) * rcp_pos(t);
            return;
        }
    }
}

// (d f / d xa) / f ;  d/d xb is its negative for every stock (stationary) leaf.
GOGP_HD double factor_dlog_xa(const DevFactor& f, double xa, double xb) {
    double r = xa - xb;
    double sg = (r > 0) - (r < 0);
    switch (f.kind) {
        case F_PARAM:
        case F_EVENTS:  // piecewise constant in the inputs
            return 0.0;
        case F_NORMAL: {
            double d = r * f.i0;
            return -d * f.i0;
        }
        case F_PERIODIC: {
            double u = fabs(r) * f.i1;
            double s, c;
            sincos(u, &s, &c);
            double d = s * f.i0;
            return -4 * d * c * sg * f.i0 * f.i1;
        }
        case F_MATERN32: {
            double d = fabs(r) * f.i0;
            return -3 * d * sg * f.i0 / (1 + GOGP_SQRT3 * d);
        }
        default: {
            double d = fabs(r) * f.i0;
            return -d * (5 - 2 * f.c + GOGP_SQRT5 * f.c * d) * sg * f.i0 / (1 + GOGP_SQRT5 * d + f.c * d * d);
        }
    }
}

#if defined(__CUDACC__)
// ---- TMA (bulk async copy) + mbarrier helpers -----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// 1-D bulk copy global -> shared through the TMA unit (SASS: UBLKCP).
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

#endif  // __CUDACC__

// Lower-triangular tile index b -> (ti, tj), ti >= tj, rows ascending.
GOGP_HD void lower_tile(int b, int& ti, int& tj) {
    int t = (int)((sqrt(8.0 * (double)b + 1.0) - 1.0) * 0.5);
    while (t * (t + 1) / 2 > b) --t;
    while ((t + 1) * (t + 2) / 2 <= b) ++t;
    ti = t;
    tj = b - t * (t + 1) / 2;
}


}  // namespace gogp
