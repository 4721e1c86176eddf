// TEST INFRASTRUCTURE: a lock-step SIMT emulator, just large enough to run the kernels of
// gogp_b200/csrc/leaf_kernels.cuh on a CPU.  One OS thread per CUDA thread; the CTAs of a launch run one
// after another in block order (the kernels here either do not talk across CTAs or only ever wait for CTAs
// that started earlier, so sequential execution is a legal schedule).  Warp collectives (__shfl_*_sync,
// __syncwarp, the m8n8k4 FP64 MMA) exchange through a per-warp slot array between two warp barriers, which
// is exactly the convergence the real instructions require: a kernel that calls them under divergent
// control flow deadlocks here instead of silently working.  __syncthreads is a CTA-wide barrier.
// Shared memory: `__shared__` variables become function statics (one CTA at a time), dynamic shared
// memory is a per-launch buffer behind GOGP_DYN_SMEM.  Built with -fsanitize=thread the same run is a
// data-race check of the kernel's shared- and global-memory protocol.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace simt {

struct uint3_t {
    unsigned x = 0, y = 0, z = 0;
};

// mbarrier (PTX): arrival count + transaction bytes per phase; a phase completes when both reach zero.
struct MBar {
    std::mutex m;
    std::condition_variable cv;
    int count = 0, pending = 0;
    long long tx = 0;
    unsigned phase = 0;
    void settle() {  // caller holds m
        if (pending == 0 && tx == 0) {
            phase ^= 1u;
            pending = count;
            cv.notify_all();
        }
    }
};

struct Cta {
    int nthreads = 0, nwarps = 0;
    std::unique_ptr<std::barrier<>> cta_bar;
    std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
    std::vector<double> slot_a, slot_b;  // [nwarps][32] exchange slots of the warp collectives
    std::vector<unsigned char> dyn;
    unsigned char* dyn_aligned = nullptr;  // 1024-byte aligned (what SWIZZLE_128B tiles need)
    std::mutex table_m;
    std::map<unsigned, std::unique_ptr<MBar>> mbars;               // keyed by shared-memory offset
    std::map<int, std::unique_ptr<std::barrier<>>> named;          // bar.sync id, count
};

struct Ctx {
    uint3_t tid, bid, bdim, gdim;
    int lane = 0, warp = 0;
    Cta* cta = nullptr;
};
inline thread_local Ctx ctx;

inline void* dyn_smem() { return ctx.cta->dyn_aligned; }
inline void syncthreads() { ctx.cta->cta_bar->arrive_and_wait(); }
inline void syncwarp() { ctx.cta->warp_bar[ctx.warp]->arrive_and_wait(); }

inline double shfl(double v, int src) {
    Cta* c = ctx.cta;
    double* s = c->slot_a.data() + ctx.warp * 32;
    s[ctx.lane] = v;
    syncwarp();
    const double r = s[src & 31];
    syncwarp();
    return r;
}

// D = A (8x4, row) * B (4x8, col) + C; lane l holds a = A[l>>2][l&3], b = B[l&3][l>>2],
// c0/c1 = C[l>>2][2(l&3) + {0,1}]  (PTX mma.sync.aligned.m8n8k4.row.col.f64)
inline void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    Cta* c = ctx.cta;
    double* sa = c->slot_a.data() + ctx.warp * 32;
    double* sb = c->slot_b.data() + ctx.warp * 32;
    sa[ctx.lane] = a;
    sb[ctx.lane] = b;
    syncwarp();
    const int row = ctx.lane >> 2, col = 2 * (ctx.lane & 3);
    for (int k = 0; k < 4; ++k) {
        c0 = std::fma(sa[row * 4 + k], sb[col * 4 + k], c0);
        c1 = std::fma(sa[row * 4 + k], sb[(col + 1) * 4 + k], c1);
    }
    syncwarp();
}

// ---- what the TMA GEMM needs: mbarriers, the 2-D tensor copy with 128-byte swizzle, named barriers ----
inline MBar& mbar_at(unsigned addr) {
    Cta* c = ctx.cta;
    std::lock_guard<std::mutex> g(c->table_m);
    auto& slot = c->mbars[addr];
    if (!slot) slot = std::make_unique<MBar>();
    return *slot;
}
inline void mbar_init(unsigned addr, int count) {
    MBar& b = mbar_at(addr);
    std::lock_guard<std::mutex> g(b.m);
    b.count = b.pending = count;
    b.tx = 0;
    b.phase = 0;
}
inline void mbar_arrive(unsigned addr) {
    MBar& b = mbar_at(addr);
    std::lock_guard<std::mutex> g(b.m);
    --b.pending;
    b.settle();
}
inline void mbar_expect_tx(unsigned addr, unsigned bytes) {  // arrive.expect_tx
    MBar& b = mbar_at(addr);
    std::lock_guard<std::mutex> g(b.m);
    b.tx += bytes;
    --b.pending;
    b.settle();
}
inline void mbar_complete_tx(unsigned addr, unsigned bytes) {
    MBar& b = mbar_at(addr);
    std::lock_guard<std::mutex> g(b.m);
    b.tx -= bytes;
    b.settle();
}
// try_wait.parity in a loop: returns once the phase with the given parity has completed
inline void mbar_wait(unsigned addr, unsigned parity) {
    MBar& b = mbar_at(addr);
    std::unique_lock<std::mutex> g(b.m);
    b.cv.wait(g, [&] { return (b.phase & 1u) != (parity & 1u); });
}
inline double lds_f64(unsigned addr) { return *reinterpret_cast<const double*>(ctx.cta->dyn_aligned + addr); }
inline void named_barrier(int id, int nthreads) {
    Cta* c = ctx.cta;
    std::barrier<>* b;
    {
        std::lock_guard<std::mutex> g(c->table_m);
        auto& slot = c->named[id];
        if (!slot) slot = std::make_unique<std::barrier<>>((std::ptrdiff_t)nthreads);
        b = slot.get();
    }
    b->arrive_and_wait();
}

inline unsigned ld_acquire(const unsigned* p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline void st_release(unsigned* p, unsigned v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }

// Launch `body` (the kernel call) over grid x block threads, CTAs in block order.
inline void launch(unsigned grid, unsigned block, size_t dyn_bytes, const std::function<void()>& body) {
    for (unsigned b = 0; b < grid; ++b) {
        Cta cta;
        cta.nthreads = (int)block;
        cta.nwarps = (int)((block + 31) / 32);
        cta.cta_bar = std::make_unique<std::barrier<>>((std::ptrdiff_t)block);
        for (int w = 0; w < cta.nwarps; ++w) {
            const int n = (w + 1) * 32 <= (int)block ? 32 : (int)block - w * 32;
            cta.warp_bar.push_back(std::make_unique<std::barrier<>>((std::ptrdiff_t)n));
        }
        cta.slot_a.assign((size_t)cta.nwarps * 32, 0.0);
        cta.slot_b.assign((size_t)cta.nwarps * 32, 0.0);
        cta.dyn.assign(dyn_bytes + 2048, 0xFF);  // poison: shared memory starts undefined
        cta.dyn_aligned = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(cta.dyn.data()) + 1023) &
                                                           ~(uintptr_t)1023);
        std::vector<std::thread> th;
        th.reserve(block);
        for (unsigned t = 0; t < block; ++t)
            th.emplace_back([&, t, b] {
                ctx.tid.x = t;
                ctx.bid.x = b;
                ctx.bdim.x = block;
                ctx.gdim.x = grid;
                ctx.lane = (int)(t & 31);
                ctx.warp = (int)(t >> 5);
                ctx.cta = &cta;
                body();
            });
        for (auto& x : th) x.join();
    }
}

}  // namespace simt

// ---- the CUDA spellings the kernels use ------------------------------------------------------
struct double2 {
    double x, y;
};
inline double2 make_double2(double x, double y) { return double2{x, y}; }

#define __global__
#define __device__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static
#define __align__(n) alignas(n)
#define threadIdx (simt::ctx.tid)
#define blockIdx (simt::ctx.bid)
#define blockDim (simt::ctx.bdim)
#define gridDim (simt::ctx.gdim)
#define __syncthreads() simt::syncthreads()
#define __syncwarp() simt::syncwarp()
#define __shfl_sync(mask, v, src) simt::shfl((v), (src))
#define __shfl_xor_sync(mask, v, o) simt::shfl((v), simt::ctx.lane ^ (o))
#define __threadfence() __atomic_thread_fence(__ATOMIC_SEQ_CST)
#define __ldcg(p) (*(p))

inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
inline int atomicCAS(int* p, int cmp, int val) {
    __atomic_compare_exchange_n(p, &cmp, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return cmp;
}
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
using std::fma;
