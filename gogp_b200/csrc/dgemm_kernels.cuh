// The cp.async FP64 DMMA GEMM kernel (all tile-map modes of the blocked algebra) in a header of its own:
// dgemm.cu includes it inside its anonymous namespace and keeps the launchers and microbenchmarks; the CPU
// test tier compiles the SAME source for the host under the SIMT emulator of tests/simt/ (GOGP_SIMT_HOST),
// where a 16-byte cp.async is an immediate copy (the earliest the data can land, so a write into a stage
// that is still being read shows up as a race) and the m8n8k4 MMA is the emulator's warp collective.
// Under nvcc nothing changes (device SASS identical).
#pragma once

constexpr int BK = 16;
constexpr int PITCH = BK + 4;  // doubles

struct GemmArgs {
    double* C;
    const double* A;
    const double* B;
    double* cdiag;
    int64_t ldc, lda, ldb;
    int tm, tn;   // tiles
    int k;        // elements
    int mode;
    double alpha, beta;
    // block-cyclic tile mask (kernels.h GemmMask), in 128-row tiles per distribution block; mtb == 0: off
    int mtb = 0, mr0 = 0, mpr = 1, mc0 = 0, mpc = 1;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
#if defined(GOGP_SIMT_HOST)
    std::memcpy(smem, gmem, 16);
#else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
#endif
}
__device__ __forceinline__ void cp_async_commit() {
#if !defined(GOGP_SIMT_HOST)
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
#if !defined(GOGP_SIMT_HOST)
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
#if defined(GOGP_SIMT_HOST)
    simt::dmma_m8n8k4(c0, c1, a, b);
#else
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
#endif
}

// Warp tile is always 64 x 32 (8 x 4 DMMA fragments); the CTA tile is
// (64*WM) x (32*WN).  Two shipped shapes:
//   <2,4,4,1>  128 x 128, 256 threads, 4 stages (160 KB), 1 CTA/SM -- required when C aliases A
//              (in-place solve with a diagonal block: one CTA must own whole rows);
//   <2,2,3,2>  128 x 64, 128 threads, 3 stages (90 KB), 2 CTAs/SM: the two resident CTAs
//              synchronise independently, so one computes while the other sits at its barrier.
//   <2,2,3,4,4> 64 x 64 and <1,4,3,3,4> 32 x 128 (in place) with 32 x 32 warp tiles (FM = 4 fragment rows): the
//              LATENCY shapes.  A launch with a handful of 128 x 128 tiles (the chain links of the blocked
//              algebra at small N: panel solves, K = 128 updates) is bound by one SM's DMMA rate per tile
//              (128^3 flops = 17 us); a quarter of the tile per CTA and half of the work per warp spreads the
//              same product over 4x the SMs.
template <int WM, int WN, int STAGES, int MINB, int FM = 8>
__global__ void __launch_bounds__(WM * WN * 32, MINB) dgemm_nt_kernel(const GemmArgs g) {
    constexpr int WROWS = 8 * FM;  // rows of a warp tile
    constexpr int BM = WROWS * WM, BN = 32 * WN, NT = WM * WN * 32;
    constexpr int STAGE_DOUBLES = (BM + BN) * PITCH;
    constexpr int RATIO = BM / BN > 0 ? BM / BN : 1;  // column tiles per diagonal block (BM >= BN)
#if defined(GOGP_SIMT_HOST)
    double* smem = reinterpret_cast<double*>(simt::dyn_smem());
#else
    extern __shared__ __align__(128) double smem[];
#endif
    int ti, tj;
    if (g.mode & GEMM_LOWER) {
        int tg;
        lower_tile(blockIdx.x / RATIO, ti, tg);
        tj = tg * RATIO + blockIdx.x % RATIO;
    } else {
        // grouped rasterisation: walk GROUP_M row tiles column by column, so the CTAs resident at
        // any time share a few A row-panels and a few B column-panels through L2 instead of
        // streaming the whole B operand from HBM once per row tile
        constexpr int GROUP_M = 16;
        const int per_group = GROUP_M * g.tn;
        const int gid = blockIdx.x / per_group, rem = blockIdx.x % per_group;
        const int first = gid * GROUP_M;
        const int gsz = (g.tm - first) < GROUP_M ? (g.tm - first) : GROUP_M;
        ti = first + rem % gsz;
        tj = rem / gsz;
    }
    if (g.mtb > 0) {
        // tile of a block-cyclic local matrix that lies strictly above the global diagonal: nothing to do
        constexpr int CPB = 128 / BN, RPB = 128 / BM;  // column / row tiles per 128 (the mask counts 128-tiles)
        const int tbn = g.mtb * CPB, tbm = g.mtb * RPB;
        const int I = g.mr0 + g.mpr * (ti / tbm), J = g.mc0 + g.mpc * (tj / tbn);
        if (J > I || (J == I && (tj % tbn) * BN > (ti % tbm) * BM + BM - 1)) return;
    }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WN, wn = warp % WN;
    const int64_t row0 = (int64_t)ti * BM, col0 = (int64_t)tj * BN;
    const int k_lo = (g.mode & GEMM_KTRI) ? ti * BM : 0;
    const int nk = (g.k - k_lo) / BK;

    const double* Ag = g.A + row0 * g.lda + k_lo;
    const double* Bg = g.B + col0 * g.ldb + k_lo;

    auto load_stage = [&](int stage, int kt) {
        double* sa = smem + stage * STAGE_DOUBLES;
        double* sb = sa + BM * PITCH;
        const int64_t koff = (int64_t)kt * BK;
#pragma unroll
        for (int q = 0; q < BM * 8 / NT; ++q) {
            const int c = tid + q * NT;
            const int r = c >> 3, kc = (c & 7) * 2;
            cp_async16(sa + r * PITCH + kc, Ag + (int64_t)r * g.lda + koff + kc);
        }
#pragma unroll
        for (int q = 0; q < BN * 8 / NT; ++q) {
            const int c = tid + q * NT;
            const int r = c >> 3, kc = (c & 7) * 2;
            cp_async16(sb + r * PITCH + kc, Bg + (int64_t)r * g.ldb + koff + kc);
        }
    };

    double acc[FM][4][2];
#pragma unroll
    for (int i = 0; i < FM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nk) load_stage(s, s);
        cp_async_commit();
    }

    const int fr = lane >> 2, fk = lane & 3;
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<STAGES - 2>();
        __syncthreads();
        {
            const int nxt = kt + STAGES - 1;
            if (nxt < nk) load_stage(nxt % STAGES, nxt);
            cp_async_commit();
        }
        const double* sa = smem + (kt % STAGES) * STAGE_DOUBLES + (wm * WROWS + fr) * PITCH + fk;
        const double* sb = smem + (kt % STAGES) * STAGE_DOUBLES + BM * PITCH + (wn * 32 + fr) * PITCH + fk;
#pragma unroll
        for (int ks = 0; ks < BK / 4; ++ks) {
            double a[FM], b[4];
#pragma unroll
            for (int i = 0; i < FM; ++i) a[i] = sa[i * 8 * PITCH + ks * 4];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = sb[j * 8 * PITCH + ks * 4];
#pragma unroll
            for (int i = 0; i < FM; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    cp_async_wait<0>();

    // epilogue: thread holds C[row][col..col+1] per fragment
    double* Cb;
    int64_t ldc;
    if ((g.mode & GEMM_DIAG_OUT) && tj / RATIO == ti) {
        Cb = g.cdiag + (int64_t)ti * BM * BM + (tj % RATIO) * BN;
        ldc = BM;
    } else {
        Cb = g.C + row0 * g.ldc + col0;
        ldc = g.ldc;
    }
    const int er = wm * WROWS + fr, ec = wn * 32 + 2 * fk;
#pragma unroll
    for (int i = 0; i < FM; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double2* p = reinterpret_cast<double2*>(Cb + (int64_t)(er + i * 8) * ldc + ec + j * 8);
            double2 v;
            v.x = g.alpha * acc[i][j][0];
            v.y = g.alpha * acc[i][j][1];
            if (g.beta != 0.0) {
                const double2 c = *p;
                v.x += g.beta * c.x;
                v.y += g.beta * c.y;
            }
            *p = v;
        }
}

