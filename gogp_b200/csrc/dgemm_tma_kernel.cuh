// The TMA + mbarrier warp-specialised FP64 DMMA GEMM kernel in a header of its own: dgemm_tma.cu includes it
// inside its anonymous namespace and keeps the tensor-map encoding and the launcher; the CPU test tier compiles
// the SAME source for the host under the SIMT emulator of tests/simt/ (GOGP_SIMT_HOST), which models the
// mbarrier phases / transaction counts, the 2-D tensor copy with its 128-byte swizzle and the named barrier,
// so the ring protocol and the swizzled fragment addressing run (and are race-checked) without a GPU.
#pragma once

constexpr int BM = 128, BN = 128, BK = 16, STAGES = 6;
constexpr int CONSUMER_WARPS = 8;
constexpr int THREADS = (CONSUMER_WARPS + 1) * 32;
constexpr int OPERAND_BYTES = BM * BK * 8;           // 16 KB
constexpr int STAGE_BYTES = 2 * OPERAND_BYTES;       // A then B
constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_BYTES + 1024 /*alignment slack*/ + 2 * STAGES * 8;

struct TmaArgs {
    double* C;
    double* cdiag;
    int64_t ldc;
    int tm, tn;
    int k;
    int mode;
    double alpha, beta;
    // block-cyclic tile mask (kernels.h GemmMask); mtb == 0: off
    int mtb = 0, mr0 = 0, mpr = 1, mc0 = 0, mpc = 1;
};

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
#if defined(GOGP_SIMT_HOST)
    simt::dmma_m8n8k4(c0, c1, a, b);
#else
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
#endif
}

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
#if defined(GOGP_SIMT_HOST)
    simt::tma_load_2d_swizzle128(dst, map, c0, c1, bar);
#else
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
#endif
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
#if defined(GOGP_SIMT_HOST)
    simt::mbar_arrive(bar);
#else
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
#endif
}
__device__ __forceinline__ void mbar_expect(uint32_t bar, uint32_t bytes) {
#if defined(GOGP_SIMT_HOST)
    simt::mbar_expect_tx(bar, bytes);
#else
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
#endif
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity) {
#if defined(GOGP_SIMT_HOST)
    simt::mbar_wait(bar, parity);
#else
    uint32_t ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!ok);
#endif
}
__device__ __forceinline__ void mbar_init_u32(uint32_t bar, int count) {
#if defined(GOGP_SIMT_HOST)
    simt::mbar_init(bar, count);
#else
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
#endif
}
__device__ __forceinline__ void mbar_init_fence() {
#if !defined(GOGP_SIMT_HOST)
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
#if defined(GOGP_SIMT_HOST)
    return simt::lds_f64(addr);
#else
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
#endif
}
// named barrier 1 over the consumer warps only (the producer warp has left the kernel by then)
__device__ __forceinline__ void consumer_barrier() {
#if defined(GOGP_SIMT_HOST)
    simt::named_barrier(1, CONSUMER_WARPS * 32);
#else
    asm volatile("bar.sync 1, %0;" ::"r"(CONSUMER_WARPS * 32) : "memory");
#endif
}

__global__ void __launch_bounds__(THREADS, 1)
    dgemm_tma_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                     const TmaArgs g) {
#if defined(GOGP_SIMT_HOST)
    const uint32_t base = 0;  // shared addresses are offsets into the emulator's (1024-byte aligned) buffer
#else
    extern __shared__ unsigned char smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B wants 1024-byte aligned tiles
#endif
    const uint32_t bars = base + STAGES * STAGE_BYTES;             // full[STAGES] then empty[STAGES]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    int ti, tj;
    if (g.mode & GEMM_LOWER) {
        lower_tile(blockIdx.x, ti, tj);
    } else {
        constexpr int GROUP_M = 16;  // grouped rasterisation, as in dgemm.cu
        const int per_group = GROUP_M * g.tn;
        const int gid = blockIdx.x / per_group, rem = blockIdx.x % per_group;
        const int first = gid * GROUP_M;
        const int gsz = (g.tm - first) < GROUP_M ? (g.tm - first) : GROUP_M;
        ti = first + rem % gsz;
        tj = rem / gsz;
    }
    if (g.mtb > 0) {
        // tile of a block-cyclic local matrix that lies strictly above the global diagonal: nothing to do
        const int I = g.mr0 + g.mpr * (ti / g.mtb), J = g.mc0 + g.mpc * (tj / g.mtb);
        if (J > I || (J == I && tj % g.mtb > ti % g.mtb)) return;
    }
    const int row0 = ti * BM, col0 = tj * BN;
    const int k_lo = (g.mode & GEMM_KTRI) ? ti * BM : 0;
    const int nk = (g.k - k_lo) / BK;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init_u32(bars + 8 * s, 1);
            mbar_init_u32(bars + 8 * (STAGES + s), CONSUMER_WARPS);
        }
        mbar_init_fence();
    }
    __syncthreads();

    if (warp == CONSUMER_WARPS) {
        // ---- producer: one lane feeds the ring ----
        if (lane == 0) {
            for (int kt = 0; kt < nk; ++kt) {
                const int s = kt % STAGES;
                if (kt >= STAGES) mbar_wait_u32(bars + 8 * (STAGES + s), ((kt / STAGES) - 1) & 1);
                const uint32_t full = bars + 8 * s;
                mbar_expect(full, STAGE_BYTES);
                const uint32_t dst = base + s * STAGE_BYTES;
                tma_load_2d(dst, &mapA, k_lo + kt * BK, row0, full);
                tma_load_2d(dst + OPERAND_BYTES, &mapB, k_lo + kt * BK, col0, full);
            }
        }
        return;
    }

    // ---- consumers ----
    const int wm = warp >> 2, wn = warp & 3;  // 2 x 4 warps, warp tile 64 x 32
    const int fr = lane >> 2, fk = lane & 3;
    uint32_t koff[4];
#pragma unroll
    for (int s = 0; s < 4; ++s) koff[s] = (uint32_t)((((((fk >> 1) << 2) + s) ^ fr) << 4) | ((fk & 1) << 3));
    const uint32_t arow = (uint32_t)((wm * 64 + fr) * 128);
    const uint32_t brow = (uint32_t)(OPERAND_BYTES + (wn * 32 + fr) * 128);

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int kt = 0; kt < nk; ++kt) {
        const int s = kt % STAGES;
        mbar_wait_u32(bars + 8 * s, (kt / STAGES) & 1);
        const uint32_t st = base + s * STAGE_BYTES;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            double a[8], b[4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
                a[i] = lds_f64(st + arow + i * 8 * 128 + koff[ks]);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                b[j] = lds_f64(st + brow + j * 8 * 128 + koff[ks]);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bars + 8 * (STAGES + s));  // this warp is done with the stage
    }

    double* Cb;
    int64_t ldc;
    if ((g.mode & GEMM_DIAG_OUT) && ti == tj) {
        Cb = g.cdiag + (int64_t)ti * BM * BN;
        ldc = BN;
    } else {
        Cb = g.C + (int64_t)row0 * g.ldc + col0;
        ldc = g.ldc;
    }
    const int er = wm * 64 + fr, ec = wn * 32 + 2 * fk;
    if (g.mode & GEMM_INPLACE) {
        // C aliases A: every consumer warp must have read its last stage before anyone stores
        consumer_barrier();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
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

