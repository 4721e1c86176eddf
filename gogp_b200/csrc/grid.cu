// gp.GP across the GPUs of one box: the CUDA + NCCL backend of the block-cyclic orchestration (grid.hpp) and
// the gogp_grid_* entry points of include/gogp_b200.h.  Reference: gp.GP.Observe / Gradient
// (gp/gp.go:374-413, 418-499) for a covariance matrix that does not fit one GPU (BASELINE configs[4]).
//
// One rank = one GPU = one gogp_handle (kernel descriptor, inputs, tile algebra through the device-level entry
// points gogp_dev_*) + two streams (a high-priority queue for the block-column chain and the collectives, a
// low-priority one for the bulk of each trailing update) + one NCCL communicator.  NCCL is bound with dlopen at
// the first grid call, so the single-GPU library has no load-time dependency on it.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <nvtx3/nvToolsExt.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gogp_b200.h"
#include "grid.hpp"
#include "kernels.h"
#include "peer_bcast.hpp"

using namespace gogp;

namespace {

// ---- NCCL, bound at run time ------------------------------------------------------------------------
struct NcclApi {
    void* lib = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitRankConfig) CommInitRankConfig = nullptr;  // optional
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclBroadcast) Broadcast = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    std::string err;
};

NcclApi* nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* env = getenv("GOGP_NCCL_LIB");
        void* lib = env ? dlopen(env, RTLD_NOW | RTLD_GLOBAL) : nullptr;
        // the copy the process already has (PyTorch ships its own libnccl.so.2) wins over the system one
        if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!lib) {
            const char* e = dlerror();
            api.err = std::string("libnccl.so.2 cannot be loaded: ") + (e ? e : "?");
            return;
        }
        api.lib = lib;
#define GOGP_NCCL_SYM(name)                                                        \
    api.name = reinterpret_cast<decltype(api.name)>(dlsym(lib, "nccl" #name));    \
    if (!api.name) api.err = "libnccl lacks nccl" #name;
        GOGP_NCCL_SYM(GetUniqueId)
        GOGP_NCCL_SYM(CommInitRank)
        GOGP_NCCL_SYM(CommDestroy)
        GOGP_NCCL_SYM(Broadcast)
        GOGP_NCCL_SYM(AllReduce)
        GOGP_NCCL_SYM(GetErrorString)
        GOGP_NCCL_SYM(GetVersion)
#undef GOGP_NCCL_SYM
        api.CommInitRankConfig =
            reinterpret_cast<decltype(api.CommInitRankConfig)>(dlsym(lib, "ncclCommInitRankConfig"));
    });
    return api.err.empty() ? &api : nullptr;
}
const char* nccl_load_error() { return "NCCL is not available (libnccl.so.2 could not be bound; set GOGP_NCCL_LIB)"; }

constexpr int kNcclMaxCtas = 0;  // cap on the CTAs of a collective (see rank_create); 0: NCCL's choice, the measured best

__global__ void int_to_double_kernel(const int* src, double* dst) { *dst = (double)*src; }

// ---- the backend grid.hpp is written against ----------------------------------------------------------
struct CudaGridBackend {
    gogp_handle* h = nullptr;
    int dev = 0, rank = 0, world = 1, nts = 0, ntn = 0;
    cudaStream_t q[2] = {nullptr, nullptr};
    ncclComm_t comm = nullptr;
    std::vector<double> ts, tn;  // natural-scale parameters of the evaluation in flight
    int* dInfo = nullptr;
    double* dTmp = nullptr;       // reductions: max(block, 8) doubles
    double* dTraceScr = nullptr;  // per-tile partial sums of the trace: (local tiles) (nts + 1) doubles
    int64_t traceScrCap = 0;
    double* hPin = nullptr;       // 64 doubles + room for alpha read-back is allocated by the caller
    int64_t hPinCap = 0;
    int64_t block = 0;
    std::vector<cudaEvent_t> ring;   // ordering events
    size_t ring_next = 0;
    std::vector<cudaEvent_t> timed;  // timing events of one evaluation
    size_t timed_used = 0;
    int64_t comm_bytes = 0, dev_bytes = 0;
    gogp_status st = GOGP_OK;  // first failure sticks; the orchestration runs on (every rank makes the same calls)
    std::string err;
    PeerLink peer;                      // panel broadcasts by the copy engines over peer memory (peer_bcast.hpp)
    size_t peer_min_bytes = 1u << 20;   // smaller messages stay on NCCL

    void fail(gogp_status s, const std::string& m) {
        if (st == GOGP_OK) {
            st = s;
            err = m;
        }
    }
    void ck(cudaError_t e, const char* what) {
        if (e != cudaSuccess) fail(e == cudaErrorMemoryAllocation ? GOGP_OUT_OF_MEMORY : GOGP_CUDA_ERROR,
                                   std::string(what) + ": " + cudaGetErrorString(e));
    }
    void ckn(ncclResult_t r, const char* what) {
        if (r != ncclSuccess) fail(GOGP_NCCL_ERROR, std::string(what) + ": " + nccl()->GetErrorString(r));
    }
    void ckg(gogp_status s) {
        if (s != GOGP_OK) fail(s, gogp_last_error(h));
    }

    double* alloc(int64_t n) {
        double* p = nullptr;
        const size_t bytes = (size_t)(n > 0 ? n : 1) * sizeof(double);
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {
            ck(e, "cudaMalloc");
            cudaGetLastError();
            return nullptr;
        }
        dev_bytes += (int64_t)bytes;
        return p;
    }
    void free(double* p) {
        if (p) cudaFree(p);
    }
    void zero(double* p, int64_t n, int qi) { ck(cudaMemsetAsync(p, 0, (size_t)n * sizeof(double), q[qi]), "memset"); }
    void zero2d(double* p, int64_t ld, int64_t rows, int64_t cols, int qi) {
        ck(cudaMemset2DAsync(p, (size_t)ld * 8, 0, (size_t)cols * 8, (size_t)rows, q[qi]), "memset2d");
    }
    void copy2d(double* dst, int64_t ldd, const double* src, int64_t lds, int64_t rows, int64_t cols, int qi) {
        // the copy engines move the panels: no SM is taken from the GEMMs
        ck(cudaMemcpy2DAsync(dst, (size_t)ldd * 8, src, (size_t)lds * 8, (size_t)cols * 8, (size_t)rows,
                             cudaMemcpyDeviceToDevice, q[qi]),
           "memcpy2d");
    }
    int record(int qi) {
        if (ring.empty()) {
            ring.resize(64);
            for (auto& e : ring) ck(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "event");
        }
        const int id = (int)(ring_next++ % ring.size());
        ck(cudaEventRecord(ring[id], q[qi]), "record");
        return id;
    }
    void wait(int qi, int ev) { ck(cudaStreamWaitEvent(q[qi], ring[ev], 0), "wait"); }
    void sync(int qi) { ck(cudaStreamSynchronize(q[qi]), "sync"); }
    int tic(int qi) {
        if (timed_used == timed.size()) {
            cudaEvent_t e;
            ck(cudaEventCreate(&e), "event");
            timed.push_back(e);
        }
        ck(cudaEventRecord(timed[timed_used], q[qi]), "record");
        return (int)timed_used++;
    }
    void tic_reset() { timed_used = 0; }
    double toc(int e0, int e1) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, timed[e0], timed[e1]) != cudaSuccess) {
            cudaGetLastError();
            return 0.0;
        }
        return ms;
    }
    void d2h(double* host, const double* devp, int64_t n, int qi) {
        if (n > hPinCap) {
            if (hPin) cudaFreeHost(hPin);
            hPin = nullptr;
            hPinCap = 0;
            ck(cudaMallocHost(&hPin, (size_t)n * sizeof(double)), "cudaMallocHost");
            if (!hPin) return;
            hPinCap = n;
        }
        ck(cudaMemcpyAsync(hPin, devp, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, q[qi]), "d2h");
        ck(cudaStreamSynchronize(q[qi]), "sync");
        std::memcpy(host, hPin, (size_t)n * sizeof(double));
    }

    void cov_block(int64_t row0, int64_t rows, int64_t col0, int64_t cols, bool diagonal, double* out, int64_t ld, int qi) {
        ckg(gogp_dev_cov_block(h, ts.data(), tn.data(), row0, rows, col0, cols, diagonal ? 1 : 0, out, ld, q[qi]));
    }
    void potrf(double* A, int64_t ld, int64_t n, double* winv, int base, int qi) {
        ckg(gogp_dev_potrf(h, A, ld, n, winv, dInfo, base, q[qi]));
    }
    void trsm(double* B, int64_t ldb, int64_t m, const double* L, int64_t ldl, int64_t n, const double* winv, int qi) {
        ckg(gogp_dev_trsm(h, B, ldb, m, L, ldl, n, winv, q[qi]));
    }
    void trtri_t(const double* L, int64_t ld, int64_t n, const double* winv, double* out, int qi) {
        ckg(gogp_dev_trtri_t(h, L, ld, n, winv, out, q[qi]));
    }
    void gemm(double* C, int64_t ldc, const double* A, int64_t lda, const double* B, int64_t ldb, int64_t m, int64_t n,
              int64_t k, double alpha, double beta, const BcMask* mk, int flags, int qi) {
        ckg(gogp_dev_gemm_bc(h, C, ldc, A, lda, B, ldb, m, n, k, alpha, beta, mk ? mk->tb : 0, mk ? mk->r0 : 0,
                             mk ? mk->pr : 1, mk ? mk->c0 : 0, mk ? mk->pc : 1, (flags & GF_KTRI) ? 1 : 0, q[qi]));
    }
    void sumlogdiag_add(const double* L, int64_t ld, int64_t nvalid, double* accp, int qi) {
        ckg(gogp_dev_sumlogdiag(h, L, ld, nvalid, dTmp, q[qi]));
        launch_axpy(accp, dTmp, 1.0, 1, q[qi]);
    }
    void gemv_acc(const double* B, int64_t ld, int64_t rows, int64_t cols, const double* v, double* accp, double a,
                  int qi) {
        launch_row_reduce(B, ld, rows, cols, v, dTmp, q[qi]);
        launch_axpy(accp, dTmp, a, rows, q[qi]);
    }
    void axpy(double* yv, const double* x, double a, int64_t n, int qi) { launch_axpy(yv, x, a, n, q[qi]); }
    void dot_add(const double* x, const double* yv, int64_t n, double* accp, int qi) {
        launch_row_reduce(x, n, 1, n, yv, dTmp, q[qi]);
        launch_axpy(accp, dTmp, 1.0, 1, q[qi]);
    }
    void trsv(const double* L, int64_t ld, const double* winv, double* rhs, double* zv, int64_t n, int qi) {
        ckg(gogp_dev_trsv(h, L, ld, winv, rhs, zv, n, q[qi]));
    }
    void trace_block(const double* alpha, const double* kinv, int64_t ld, int64_t row0, int64_t rows, int64_t col0,
                     int64_t cols, double* accp, int qi) {
        ckg(gogp_dev_trace_block(h, ts.data(), alpha, kinv, ld, row0, rows, col0, cols, accp, dTraceScr, q[qi]));
    }
    void trace_local(const double* alpha, const double* kinv, int64_t ld, int64_t rows, int64_t cols, const BcMask& mk,
                     double* accp, int qi) {
        const int64_t need = (rows / TILE) * (cols / TILE) * (nts + 1);
        if (need > traceScrCap) {  // first evaluation after set_data: before any collective of the trace phase
            free(dTraceScr);
            dTraceScr = alloc(need);
            traceScrCap = dTraceScr ? need : 0;
            if (!dTraceScr) return;
        }
        ckg(gogp_dev_trace_local(h, ts.data(), alpha, kinv, ld, rows, cols, mk.tb, mk.r0, mk.pr, mk.c0, mk.pc, accp,
                                 dTraceScr, q[qi]));
    }
    void range_push(const char* name) { nvtxRangePushA(name); }  // NVTX: free unless a profiler is attached
    void range_pop() { nvtxRangePop(); }
    void info_reset(int qi) { ck(cudaMemsetAsync(dInfo, 0, sizeof(int), q[qi]), "memset"); }
    void info_to(double* dst, int qi) { int_to_double_kernel<<<1, 1, 0, q[qi]>>>(dInfo, dst); }
    int info_host() {
        int v = 0;
        ck(cudaMemcpy(&v, dInfo, sizeof(int), cudaMemcpyDeviceToHost), "memcpy");
        return v;
    }

    // ---- collectives (NCCL over NVLink), on the priority queue ---------------------------------------
    void bcast(double* p, int64_t n, int root, int qi) {
        if (world == 1 || n <= 0) return;
        if (peer.usable(p, (size_t)n * 8, peer_min_bytes)) {
            // every rank takes this branch together: the buffers, sizes and offsets are the same everywhere
            if (!peer.bcast(p, (size_t)n * 8, root, q[qi])) fail(GOGP_NCCL_ERROR, peer.err);
            if (rank != root) comm_bytes += 8 * n;
            return;
        }
        ckn(nccl()->Broadcast(p, p, (size_t)n, ncclDouble, root, comm, q[qi]), "ncclBroadcast");
        if (rank != root) comm_bytes += 8 * n;
    }
    void allreduce_sum(double* p, int64_t n, int qi) {
        if (world == 1 || n <= 0) return;
        ckn(nccl()->AllReduce(p, p, (size_t)n, ncclDouble, ncclSum, comm, q[qi]), "ncclAllReduce");
        comm_bytes += 8 * n;
    }
    void allreduce_max(double* p, int64_t n, int qi) {
        if (world == 1 || n <= 0) return;
        ckn(nccl()->AllReduce(p, p, (size_t)n, ncclDouble, ncclMax, comm, q[qi]), "ncclAllReduce");
        comm_bytes += 8 * n;
    }
};

struct GridRank {
    CudaGridBackend be;
    BlockCyclic<CudaGridBackend> bc{be};
    bool inited = false;
    int64_t N = 0;
    bool observed = false;
    double lml = 0.0;
    std::vector<double> grad;
    gogp_status st = GOGP_OK;
};

}  // namespace

struct gogp_grid {
    std::vector<std::unique_ptr<GridRank>> ranks;  // the ranks this process drives
    std::unique_ptr<PeerCtl> peer_ctl;             // rendezvous counters of the peer broadcast when the ranks are threads
    int ndim = 1, nts = 0, ntn = 0, world = 1, pr = 1, pc = 1;
    int64_t block = 2048;
    std::string err;
};

namespace {

void default_grid(int world, int* pr, int* pc) {  // pr >= pc, powers of two while they divide
    int c = 1;
    while ((c * 2) * (c * 2) <= world && world % (c * 2) == 0) c *= 2;
    *pc = c;
    *pr = world / c;
}

gogp_status gfail(gogp_grid* g, gogp_status s, const std::string& m) {
    g->err = m;
    return s;
}

// run f on every local rank (a host thread per device when there are several) and fold the statuses
template <class F>
gogp_status run_all(gogp_grid* g, F f) {
    const size_t n = g->ranks.size();
    std::vector<gogp_status> st(n, GOGP_OK);
    if (n == 1) {
        st[0] = f(*g->ranks[0]);
    } else {
        std::vector<std::thread> th;
        for (size_t i = 0; i < n; ++i) th.emplace_back([&, i] { st[i] = f(*g->ranks[i]); });
        for (auto& t : th) t.join();
    }
    for (size_t i = 0; i < n; ++i)
        if (st[i] != GOGP_OK) {
            g->err = g->ranks[i]->be.err.empty() ? gogp_status_string(st[i]) : g->ranks[i]->be.err;
            return st[i];
        }
    return GOGP_OK;
}

gogp_status rank_create(GridRank& r, int ndim, const gogp_op* simil, int n_simil_ops, int nts, const gogp_op* noise,
                        int n_noise_ops, int ntn, int device, int rank, int world, int64_t block,
                        const unsigned char* id, PeerCtl* shared_ctl) {
    CudaGridBackend& be = r.be;
    be.dev = device;
    be.rank = rank;
    be.world = world;
    be.block = block;
    gogp_status s = gogp_create(ndim, simil, n_simil_ops, nts, noise, n_noise_ops, ntn, device, &be.h);
    if (s != GOGP_OK) {
        be.fail(s, be.h ? gogp_last_error(be.h) : "gogp_create failed");
        return s;
    }
    be.nts = nts;
    be.ntn = (noise && n_noise_ops > 0) ? ntn : 0;
    be.ts.assign(nts > 0 ? nts : 1, 1.0);
    be.tn.assign(be.ntn > 0 ? be.ntn : 1, 1.0);
    be.ck(cudaSetDevice(device), "cudaSetDevice");
    int lo = 0, hi = 0;
    be.ck(cudaDeviceGetStreamPriorityRange(&lo, &hi), "priority range");
    be.ck(cudaStreamCreateWithPriority(&be.q[GQ_MAIN], cudaStreamNonBlocking, hi), "stream");
    be.ck(cudaStreamCreateWithPriority(&be.q[GQ_SIDE], cudaStreamNonBlocking, lo), "stream");
    be.ck(cudaMalloc(&be.dInfo, sizeof(int)), "cudaMalloc");
    const int64_t tmpn = block > 8 ? block : 8;
    be.dTmp = be.alloc(tmpn);
    be.traceScrCap = (block / TILE) * (block / TILE) * (nts + 1);
    be.dTraceScr = be.alloc(be.traceScrCap);
    if (be.st != GOGP_OK) return be.st;
    if (world > 1) {
        NcclApi* api = nccl();
        if (!api) {
            be.fail(GOGP_NCCL_ERROR, nccl_load_error());
            return be.st;
        }
        ncclUniqueId uid;
        static_assert(sizeof(uid) == GOGP_GRID_ID_BYTES, "ncclUniqueId is 128 bytes");
        std::memcpy(&uid, id, sizeof(uid));
        // The collectives share the GPU with the DMMA GEMMs of the side queue: every CTA NCCL holds is an SM the trailing
        // update cannot use (a 197 KB GEMM CTA shares its SM with nobody), but fewer CTAs make a slower broadcast on the
        // block-column chain.  Measured on 4 GPUs at N = 65536, maxCTAs = 1 / 2 / 4 / 8 / NCCL's choice: 3477 / 2588 /
        // 2371 / 2304 / 2310 ms per evaluation -- so the default leaves the choice to NCCL (GOGP_NCCL_MAX_CTAS = 0); the
        // way to get the SMs back is the copy-engine broadcast below, not a cap.
        static const int max_ctas = getenv("GOGP_NCCL_MAX_CTAS") ? atoi(getenv("GOGP_NCCL_MAX_CTAS")) : kNcclMaxCtas;
        if (max_ctas > 0 && api->CommInitRankConfig) {
            ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
            cfg.minCTAs = 1;
            cfg.maxCTAs = max_ctas;
            be.ckn(api->CommInitRankConfig(&be.comm, world, uid, rank, &cfg), "ncclCommInitRankConfig");
        } else {
            be.ckn(api->CommInitRank(&be.comm, world, uid, rank), "ncclCommInitRank");
        }
        // GOGP_PEER_BCAST=1: the panel broadcasts go over peer memory with the copy engines when every rank can map
        // every other (peer_bcast.hpp).  Opt-in: verified on two GPUs (threads and processes) only; NCCL otherwise.
        static const int peer_on = getenv("GOGP_PEER_BCAST") ? atoi(getenv("GOGP_PEER_BCAST")) : 0;
        if (be.st == GOGP_OK && peer_on && be.peer.attach(shared_ctl, id, rank, world, device)) be.peer.setup_events();
    }
    return be.st;
}

// Teardown in two phases.  First, in parallel on every local rank: drain the streams and take the peer mappings down
// (a rendezvous of the ranks).  Then, one rank after the other: free memory and destroy the communicator -- cudaFree
// synchronises its device, and doing that on one thread while another is inside ncclCommDestroy is the kind of
// concurrency NCCL warns against.
void rank_quiesce(GridRank& r) {
    CudaGridBackend& be = r.be;
    cudaSetDevice(be.dev);
    if (be.q[GQ_MAIN]) cudaStreamSynchronize(be.q[GQ_MAIN]);
    if (be.q[GQ_SIDE]) cudaStreamSynchronize(be.q[GQ_SIDE]);
    be.peer.unregister_memory();
    be.peer.detach();
}

void rank_destroy(GridRank& r) {
    CudaGridBackend& be = r.be;
    cudaSetDevice(be.dev);
    if (r.inited) r.bc.release();
    if (be.comm && nccl()) nccl()->CommDestroy(be.comm);
    be.free(be.dTmp);
    be.free(be.dTraceScr);
    if (be.dInfo) cudaFree(be.dInfo);
    if (be.hPin) cudaFreeHost(be.hPin);
    for (auto& e : be.ring) cudaEventDestroy(e);
    for (auto& e : be.timed) cudaEventDestroy(e);
    if (be.q[GQ_MAIN]) cudaStreamDestroy(be.q[GQ_MAIN]);
    if (be.q[GQ_SIDE]) cudaStreamDestroy(be.q[GQ_SIDE]);
    if (be.h) gogp_destroy(be.h);
}

gogp_status check_shape(int world, int* pr, int* pc, int64_t* block) {
    if (world < 1) return GOGP_BAD_ARGUMENT;
    if (*pr == 0 && *pc == 0) default_grid(world, pr, pc);
    if (*pr < 1 || *pc < 1 || *pr * *pc != world) return GOGP_BAD_ARGUMENT;
    if (*block == 0) *block = 2048;
    if (*block < TILE || *block % TILE) return GOGP_BAD_ARGUMENT;
    return GOGP_OK;
}

// every rank agrees on a status (a rank that failed locally must not leave the others inside a collective)
gogp_status agree(GridRank& r, gogp_status mine) {
    CudaGridBackend& be = r.be;
    if (be.world == 1 || !be.comm) return mine;
    double v = (double)(int)mine;
    cudaSetDevice(be.dev);
    if (cudaMemcpyAsync(be.dTmp, &v, sizeof(double), cudaMemcpyHostToDevice, be.q[GQ_MAIN]) != cudaSuccess) return mine;
    if (nccl()->AllReduce(be.dTmp, be.dTmp, 1, ncclDouble, ncclMax, be.comm, be.q[GQ_MAIN]) != ncclSuccess)
        return mine != GOGP_OK ? mine : GOGP_NCCL_ERROR;
    if (cudaMemcpyAsync(&v, be.dTmp, sizeof(double), cudaMemcpyDeviceToHost, be.q[GQ_MAIN]) != cudaSuccess) return mine;
    cudaStreamSynchronize(be.q[GQ_MAIN]);
    const gogp_status all = (gogp_status)(int)v;
    if (mine == GOGP_OK && all != GOGP_OK) be.fail(all, "another rank of the grid failed");
    return mine != GOGP_OK ? mine : all;
}

gogp_status rank_set_data(gogp_grid* g, GridRank& r, const double* X, const double* Y, int64_t N) {
    CudaGridBackend& be = r.be;
    be.st = GOGP_OK;
    be.err.clear();
    cudaSetDevice(be.dev);
    if (r.inited) {
        cudaStreamSynchronize(be.q[GQ_MAIN]);
        cudaStreamSynchronize(be.q[GQ_SIDE]);
        be.peer.unregister_memory();
        r.bc.release();
        r.inited = false;
    }
    r.observed = false;
    r.N = N;
    gogp_status mine = GOGP_OK;
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    const int64_t need = BlockCyclic<CudaGridBackend>::bytes_needed(N, g->block, g->pr, g->pc) + (int64_t)N * g->block * 8;
    if ((int64_t)free_b < need) {
        char buf[200];
        snprintf(buf, sizeof buf, "rank %d needs %.1f GB of device memory for N = %lld on a %d x %d grid, %.1f GB free",
                 be.rank, need / 1e9, (long long)N, g->pr, g->pc, free_b / 1e9);
        be.fail(GOGP_OUT_OF_MEMORY, buf);
        mine = GOGP_OUT_OF_MEMORY;
    }
    if (mine == GOGP_OK) {
        be.dev_bytes = 0;
        be.ckg(gogp_dev_set_inputs(be.h, X, N, g->block));
        if (be.st == GOGP_OK && !r.bc.init(N, g->block, be.rank, be.world, g->pr, g->pc, be.nts)) {
            be.fail(GOGP_OUT_OF_MEMORY, "out of device memory for the rank's share of K");
            r.bc.release();
        } else if (be.st == GOGP_OK) {
            r.inited = true;
            // the panel solves of one step cover at most the rank's whole block column
            be.ckg(gogp_dev_reserve(be.h, (int64_t)(r.bc.nr > 0 ? r.bc.nr : 1) * g->block));
            be.zero(r.bc.y, r.bc.Npad, GQ_MAIN);
            be.ck(cudaMemcpyAsync(r.bc.y, Y, (size_t)N * sizeof(double), cudaMemcpyHostToDevice, be.q[GQ_MAIN]), "h2d");
            be.sync(GQ_MAIN);
        }
        mine = be.st;
    }
    const gogp_status all = agree(r, mine);
    if (all != GOGP_OK && r.inited) {
        r.bc.release();
        r.inited = false;
    }
    if (all == GOGP_OK && be.world > 1 && be.peer.events_ok) {
        // the buffers the big broadcasts live in: the two panel generations and the diagonal-block staging
        void* ptrs[3] = {r.bc.panel[0], r.bc.panel[1], r.bc.dk};
        const size_t sizes[3] = {(size_t)r.bc.nb * r.bc.bsz * 8, (size_t)r.bc.nb * r.bc.bsz * 8, (size_t)r.bc.dk_len() * 8};
        be.peer.register_memory(ptrs, sizes, 3);
    }
    return all;
}

gogp_status rank_observe(gogp_grid* g, GridRank& r, const double* theta_s, const double* theta_n) {
    CudaGridBackend& be = r.be;
    if (!r.inited) {
        be.fail(GOGP_NOT_READY, "gogp_grid_observe before gogp_grid_set_data");
        return GOGP_NOT_READY;
    }
    be.st = GOGP_OK;
    be.err.clear();
    cudaSetDevice(be.dev);
    for (int i = 0; i < g->nts; ++i) be.ts[i] = theta_s[i];
    for (int i = 0; i < g->ntn; ++i) be.tn[i] = theta_n[i];
    r.observed = false;
    int pivot = 0;
    double lml = 0.0;
    const bool ok = r.bc.observe(&lml, &pivot);
    if (be.st != GOGP_OK) return be.st;
    if (!ok) {
        char buf[160];
        if (pivot > 0)
            snprintf(buf, sizeof buf, "Factorize: covariance matrix is not positive definite (pivot %d of %lld)", pivot,
                     (long long)r.N);
        else
            snprintf(buf, sizeof buf, "Factorize: covariance matrix is not positive definite");
        be.fail(GOGP_NOT_POSITIVE_DEFINITE, buf);
        return be.st;
    }
    r.lml = lml;
    r.observed = true;
    return GOGP_OK;
}

gogp_status rank_gradient(gogp_grid* g, GridRank& r) {
    CudaGridBackend& be = r.be;
    if (!r.inited || !r.observed) {
        be.fail(GOGP_NOT_READY, "gogp_grid_gradient before gogp_grid_observe");
        return GOGP_NOT_READY;
    }
    be.st = GOGP_OK;
    be.err.clear();
    cudaSetDevice(be.dev);
    std::vector<double> sums(g->nts + 1, 0.0);
    r.bc.gradient(sums.data());
    if (be.st != GOGP_OK) return be.st;
    r.grad.assign(g->nts + g->ntn, 0.0);
    for (int q = 0; q < g->nts; ++q) r.grad[q] = sums[q];
    if (g->ntn > 0) {
        double var = 0.0;
        std::vector<double> dlog(g->ntn, 0.0);
        be.ckg(gogp_noise_eval(be.h, be.tn.data(), &var, dlog.data()));
        // every shipped noise is input-independent (kernel/noise.go:21-53): d LML / d log theta_n = 0.5 tr(W) d var
        for (int q = 0; q < g->ntn; ++q) r.grad[g->nts + q] = 0.5 * sums[g->nts] * dlog[q];
    }
    return be.st;
}

}  // namespace

extern "C" {

gogp_status gogp_grid_unique_id(unsigned char id[GOGP_GRID_ID_BYTES]) {
    if (!id) return GOGP_BAD_ARGUMENT;
    NcclApi* api = nccl();
    if (!api) return GOGP_NCCL_ERROR;
    ncclUniqueId uid;
    if (api->GetUniqueId(&uid) != ncclSuccess) return GOGP_NCCL_ERROR;
    std::memcpy(id, &uid, sizeof(uid));
    return GOGP_OK;
}

gogp_status gogp_grid_create_rank(int ndim, const gogp_op* simil, int n_simil_ops, int ntheta_simil, const gogp_op* noise,
                                  int n_noise_ops, int ntheta_noise, int device, int rank, int world, int pr, int pc,
                                  int64_t block, const unsigned char id[GOGP_GRID_ID_BYTES], gogp_grid** out) {
    if (!out) return GOGP_BAD_ARGUMENT;
    gogp_grid* g = new gogp_grid();
    *out = g;  // returned even on failure so gogp_grid_last_error can be read; the caller destroys it
    if (check_shape(world, &pr, &pc, &block) != GOGP_OK || rank < 0 || rank >= world || (world > 1 && !id))
        return gfail(g, GOGP_BAD_ARGUMENT, "bad grid shape: pr * pc must equal the number of ranks, block a multiple of 128");
    g->ndim = ndim;
    g->nts = ntheta_simil;
    g->ntn = (noise && n_noise_ops > 0) ? ntheta_noise : 0;
    g->world = world;
    g->pr = pr;
    g->pc = pc;
    g->block = block;
    g->ranks.emplace_back(new GridRank());
    const gogp_status s = rank_create(*g->ranks[0], ndim, simil, n_simil_ops, ntheta_simil, noise, n_noise_ops,
                                      ntheta_noise, device, rank, world, block, id, nullptr);
    if (s != GOGP_OK) g->err = g->ranks[0]->be.err;
    return s;
}

gogp_status gogp_create_grid(int ndim, const gogp_op* simil, int n_simil_ops, int ntheta_simil, const gogp_op* noise,
                             int n_noise_ops, int ntheta_noise, const int* devices, int ndev, int pr, int pc,
                             int64_t block, gogp_grid** out) {
    if (!out) return GOGP_BAD_ARGUMENT;
    gogp_grid* g = new gogp_grid();
    *out = g;
    if (!devices || check_shape(ndev, &pr, &pc, &block) != GOGP_OK)
        return gfail(g, GOGP_BAD_ARGUMENT, "bad grid shape: pr * pc must equal the number of devices, block a multiple of 128");
    g->ndim = ndim;
    g->nts = ntheta_simil;
    g->ntn = (noise && n_noise_ops > 0) ? ntheta_noise : 0;
    g->world = ndev;
    g->pr = pr;
    g->pc = pc;
    g->block = block;
    unsigned char id[GOGP_GRID_ID_BYTES] = {0};
    if (ndev > 1) {
        if (gogp_grid_unique_id(id) != GOGP_OK) return gfail(g, GOGP_NCCL_ERROR, nccl_load_error());
    }
    for (int i = 0; i < ndev; ++i) g->ranks.emplace_back(new GridRank());
    g->peer_ctl.reset(new PeerCtl());  // value-initialised: all counters zero
    // ncclCommInitRank blocks until every rank has joined: one host thread per device
    std::vector<int> devs(devices, devices + ndev);
    return run_all(g, [&](GridRank& r) {
        int pos = 0;  // the rank of a GridRank is its position in g->ranks
        for (size_t k = 0; k < g->ranks.size(); ++k)
            if (g->ranks[k].get() == &r) pos = (int)k;
        return rank_create(r, ndim, simil, n_simil_ops, ntheta_simil, noise, n_noise_ops, ntheta_noise, devs[pos], pos, ndev,
                           block, id, g->peer_ctl.get());
    });
}

void gogp_grid_destroy(gogp_grid* g) {
    if (!g) return;
    run_all(g, [&](GridRank& r) {
        rank_quiesce(r);
        return GOGP_OK;
    });
    for (auto& r : g->ranks) rank_destroy(*r);
    delete g;
}

gogp_status gogp_grid_set_data(gogp_grid* g, const double* X, const double* Y, int64_t N) {
    if (!g || g->ranks.empty()) return GOGP_BAD_ARGUMENT;
    if (!X || !Y || N <= 0) return gfail(g, GOGP_BAD_ARGUMENT, "X, Y and N > 0 are required");
    return run_all(g, [&](GridRank& r) { return rank_set_data(g, r, X, Y, N); });
}

gogp_status gogp_grid_absorb(gogp_grid* g, const double* theta_simil, const double* theta_noise) {
    if (!g || g->ranks.empty()) return GOGP_BAD_ARGUMENT;
    if ((!theta_simil && g->nts > 0) || (!theta_noise && g->ntn > 0)) return gfail(g, GOGP_BAD_ARGUMENT, "theta is NULL");
    return run_all(g, [&](GridRank& r) { return rank_observe(g, r, theta_simil, theta_noise); });
}

gogp_status gogp_grid_observe(gogp_grid* g, const double* log_theta, double* lml) {
    if (!g || g->ranks.empty() || !lml) return GOGP_BAD_ARGUMENT;
    if (!log_theta && g->nts + g->ntn > 0) return gfail(g, GOGP_BAD_ARGUMENT, "log_theta is NULL");
    std::vector<double> ts(g->nts > 0 ? g->nts : 1), tn(g->ntn > 0 ? g->ntn : 1);
    for (int i = 0; i < g->nts; ++i) ts[i] = std::exp(log_theta[i]);  // gp/gp.go:378-381
    for (int i = 0; i < g->ntn; ++i) tn[i] = std::exp(log_theta[g->nts + i]);
    const gogp_status s = gogp_grid_absorb(g, ts.data(), tn.data());
    if (s != GOGP_OK) return s;
    *lml = g->ranks[0]->lml;
    return GOGP_OK;
}

gogp_status gogp_grid_lml(gogp_grid* g, double* lml) {
    if (!g || g->ranks.empty() || !lml) return GOGP_BAD_ARGUMENT;
    if (!g->ranks[0]->observed) return gfail(g, GOGP_NOT_READY, "LML before Observe/Absorb");
    *lml = g->ranks[0]->lml;
    return GOGP_OK;
}

gogp_status gogp_grid_gradient(gogp_grid* g, double* grad, int64_t len) {
    if (!g || g->ranks.empty() || (!grad && len > 0)) return GOGP_BAD_ARGUMENT;
    if (len != g->nts + g->ntn) return gfail(g, GOGP_BAD_ARGUMENT, "gradient length must be ntheta_simil + ntheta_noise");
    const gogp_status s = run_all(g, [&](GridRank& r) { return rank_gradient(g, r); });
    if (s != GOGP_OK) return s;
    for (int64_t i = 0; i < len; ++i) grad[i] = g->ranks[0]->grad[i];
    return GOGP_OK;
}

gogp_status gogp_grid_get_alpha(gogp_grid* g, double* alpha, int64_t N) {
    if (!g || g->ranks.empty() || !alpha) return GOGP_BAD_ARGUMENT;
    GridRank& r = *g->ranks[0];
    if (!r.inited || !r.bc.have_kinv || N != r.N) return gfail(g, GOGP_NOT_READY, "alpha exists after gogp_grid_gradient");
    cudaSetDevice(r.be.dev);
    r.be.st = GOGP_OK;
    r.be.d2h(alpha, r.bc.alpha, N, GQ_MAIN);
    if (r.be.st != GOGP_OK) g->err = r.be.err;
    return r.be.st;
}

gogp_status gogp_grid_phase_times(const gogp_grid* g, double* ms, double* comm_ms) {
    if (!g || g->ranks.empty() || !ms) return GOGP_BAD_ARGUMENT;
    for (int p = 0; p < GOGP_GRID_NPHASE; ++p) {
        ms[p] = 0.0;
        if (comm_ms) comm_ms[p] = 0.0;
        for (const auto& r : g->ranks) {
            if (r->bc.phase_ms[p] > ms[p]) ms[p] = r->bc.phase_ms[p];
            if (comm_ms && r->bc.comm_ms[p] > comm_ms[p]) comm_ms[p] = r->bc.comm_ms[p];
        }
    }
    return GOGP_OK;
}

gogp_status gogp_grid_stats(const gogp_grid* g, double* stats) {
    if (!g || g->ranks.empty() || !stats) return GOGP_BAD_ARGUMENT;
    const GridRank& r = *g->ranks[0];
    int ver = 0;
    if (g->world > 1 && nccl()) nccl()->GetVersion(&ver);
    stats[0] = (double)r.be.comm_bytes;
    stats[1] = (double)gogp_launch_count(r.be.h);
    stats[2] = g->pr;
    stats[3] = g->pc;
    stats[4] = (double)g->block;
    stats[5] = (double)r.be.dev_bytes;
    stats[6] = (double)ver;
    stats[7] = (double)g->world;
    stats[8] = (double)r.be.peer.pulled_bytes;
    // device time of the last Observe + Gradient: the slowest local rank's own sum of phases (the per-phase maxima of
    // gogp_grid_phase_times do not add up when the ranks drift apart inside an evaluation)
    double worst = 0.0;
    for (const auto& rk : g->ranks) {
        double sum = 0.0;
        for (int p = 0; p < GOGP_GRID_NPHASE; ++p) sum += rk->bc.phase_ms[p];
        if (sum > worst) worst = sum;
    }
    stats[9] = worst;
    return GOGP_OK;
}

const char* gogp_grid_last_error(const gogp_grid* g) { return g ? g->err.c_str() : "null grid"; }

}  // extern "C"
