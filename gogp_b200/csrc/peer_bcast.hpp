// Panel broadcast over NVLink peer memory, driven by the copy engines.
//
// The block-cyclic factorisation (grid.hpp) moves a whole panel (up to N x 2048 doubles) from the ranks of one
// process column to every rank, every step, while the DMMA GEMMs of the trailing update run.  An NCCL broadcast
// does that with CTAs: each one takes an SM away from the GEMM (whose 197 KB CTAs cannot share one) for as long
// as the collective is in flight, most of it waiting for a peer.  Here the receivers PULL the root's buffer with
// cudaMemcpyAsync over a peer mapping: the transfer runs on a copy engine at NVLink rate and no SM is touched.
//
//   root      record `ready` on its stream  -> post (host counter) -> wait (host) for every receiver's ack post,
//             make its stream wait for their `ack` events (the buffer may be overwritten afterwards)
//   receiver  wait (host) for the root's post -> make its stream wait for `ready` -> cudaMemcpyAsync(mine <- root's,
//             peer) -> record `ack` -> post
//
// The host counters exist because cudaStreamWaitEvent binds to the record that has been CALLED by then: a
// receiver must not call it before the root has called its record.  Every rank makes the same calls in the same
// order (SPMD), so the (root, sequence number) of a broadcast identifies it.
//
// Two transports behind one protocol: the ranks are host threads of one process (gogp_create_grid: raw pointers and
// events, cudaDeviceEnablePeerAccess) or one process each (gogp_grid_create_rank under torchrun: CUDA IPC memory and
// event handles, the counters in a POSIX shared-memory segment named after the grid's unique id).  Small messages
// and anything outside the registered buffers stay on NCCL; so does everything when a peer mapping is unavailable.
#pragma once
#include <cuda_runtime.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>

namespace gogp {

constexpr int kPeerMaxRanks = 16, kPeerEv = 4, kPeerMaxAlloc = 4;
constexpr uint64_t kPeerMagic = 0x676f67705f623230ull;

// Zero-initialised; lives on the heap (threads) or in POSIX shared memory (processes).
struct PeerCtl {
    std::atomic<uint64_t> magic;
    std::atomic<uint64_t> posted[kPeerMaxRanks];                // broadcasts rank r has posted as root
    std::atomic<uint64_t> acked[kPeerMaxRanks][kPeerMaxRanks];  // [receiver][root]: pulls the receiver has queued
    std::atomic<uint32_t> bar_count, bar_gen;
    std::atomic<int> ok[kPeerMaxRanks];
    int dev[kPeerMaxRanks];
    // processes: handles
    cudaIpcMemHandle_t mem_h[kPeerMaxRanks][kPeerMaxAlloc];
    cudaIpcEventHandle_t ready_h[kPeerMaxRanks][kPeerEv];
    cudaIpcEventHandle_t ack_h[kPeerMaxRanks][kPeerMaxRanks][kPeerEv];
    // threads: the objects themselves
    void* base[kPeerMaxRanks][kPeerMaxAlloc];
    cudaEvent_t ready_e[kPeerMaxRanks][kPeerEv];
    cudaEvent_t ack_e[kPeerMaxRanks][kPeerMaxRanks][kPeerEv];
};

struct PeerLink {
    PeerCtl* ctl = nullptr;
    bool ipc = false;
    int rank = 0, world = 1, dev = 0;
    bool events_ok = false, mem_ok = false;
    cudaEvent_t ready[kPeerMaxRanks][kPeerEv] = {};
    cudaEvent_t ack[kPeerMaxRanks][kPeerMaxRanks][kPeerEv] = {};  // [receiver][root]
    char* base[kPeerMaxRanks][kPeerMaxAlloc] = {};
    size_t bytes[kPeerMaxAlloc] = {};
    int nalloc = 0;
    uint64_t nroot[kPeerMaxRanks] = {};
    double timeout_s = 30.0;
    int64_t pulled_bytes = 0;
    bool broken = false;  // a rendezvous failed: every later broadcast fails at once instead of timing out again
    int debug = 0;
    std::string err;

    // ---- host rendezvous ----------------------------------------------------------------------
    template <class A, class V>
    bool wait_ge(A& a, V v, double limit_s) {
        const auto t0 = std::chrono::steady_clock::now();
        for (uint64_t spin = 0;; ++spin) {
            if (a.load(std::memory_order_acquire) >= v) return true;
            if ((spin & 0x3ff) == 0x3ff) {
                std::this_thread::yield();
                if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > limit_s) return false;
            }
        }
    }
    bool barrier(double limit_s) {
        const uint32_t g = ctl->bar_gen.load(std::memory_order_acquire);
        if (ctl->bar_count.fetch_add(1, std::memory_order_acq_rel) + 1 == (uint32_t)world) {
            ctl->bar_count.store(0, std::memory_order_relaxed);
            ctl->bar_gen.store(g + 1, std::memory_order_release);
            return true;
        }
        const auto t0 = std::chrono::steady_clock::now();
        for (uint64_t spin = 0;; ++spin) {
            if (ctl->bar_gen.load(std::memory_order_acquire) != g) return true;
            if ((spin & 0x3ff) == 0x3ff) {
                std::this_thread::yield();
                if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > limit_s) return false;
            }
        }
    }
    // every rank reports, every rank learns whether all succeeded
    bool all_ok(bool mine) {
        ctl->ok[rank].store(mine ? 1 : 0, std::memory_order_release);
        if (!barrier(timeout_s)) return false;
        bool all = true;
        for (int r = 0; r < world; ++r) all = all && ctl->ok[r].load(std::memory_order_acquire) == 1;
        if (!barrier(timeout_s)) return false;
        return all;
    }

    // ---- attach ----------------------------------------------------------------------------------
    // threads: `shared` is the grid's PeerCtl; processes: shared == nullptr and the segment is named after id
    bool attach(PeerCtl* shared, const unsigned char* id, int rank_, int world_, int dev_) {
        rank = rank_;
        world = world_;
        dev = dev_;
        debug = getenv("GOGP_PEER_DEBUG") ? atoi(getenv("GOGP_PEER_DEBUG")) : 0;
        if (world > kPeerMaxRanks) return false;
        if (shared) {
            ctl = shared;
            ipc = false;
        } else {
            ipc = true;
            uint64_t h = 1469598103934665603ull;
            for (int i = 0; i < 128; ++i) h = (h ^ id[i]) * 1099511628211ull;
            char name[64];
            snprintf(name, sizeof name, "/gogp_%016llx", (unsigned long long)h);
            int fd = -1;
            const auto t0 = std::chrono::steady_clock::now();
            if (rank == 0) {
                shm_unlink(name);
                fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
                if (fd < 0 || ftruncate(fd, (off_t)sizeof(PeerCtl)) != 0) {
                    if (fd >= 0) close(fd);
                    return false;
                }
            } else {
                for (;;) {
                    fd = shm_open(name, O_RDWR, 0600);
                    struct stat sb;
                    if (fd >= 0 && fstat(fd, &sb) == 0 && (size_t)sb.st_size == sizeof(PeerCtl)) break;
                    if (fd >= 0) close(fd);
                    fd = -1;
                    std::this_thread::sleep_for(std::chrono::milliseconds(2));
                    if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 60.0) return false;
                }
            }
            void* m = mmap(nullptr, sizeof(PeerCtl), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
            close(fd);
            if (m == MAP_FAILED) return false;
            ctl = static_cast<PeerCtl*>(m);
            if (rank == 0) ctl->magic.store(kPeerMagic, std::memory_order_release);
            if (!wait_ge(ctl->magic, kPeerMagic, 60.0)) return false;
            shm_name = name;
        }
        ctl->dev[rank] = dev;
        if (!barrier(60.0)) return false;
        if (ipc && rank == 0) shm_unlink(shm_name.c_str());  // everyone is attached: the name can go
        return true;
    }
    std::string shm_name;

    bool setup_events() {
        bool mine = true;
        const unsigned flags = cudaEventDisableTiming | (ipc ? cudaEventInterprocess : 0);
        for (int i = 0; i < kPeerEv; ++i) {
            mine = mine && cudaEventCreateWithFlags(&ready[rank][i], flags) == cudaSuccess;
            for (int root = 0; root < world; ++root)
                mine = mine && cudaEventCreateWithFlags(&ack[rank][root][i], flags) == cudaSuccess;
        }
        if (mine) {
            for (int i = 0; i < kPeerEv; ++i) {
                if (ipc) {
                    mine = mine && cudaIpcGetEventHandle(&ctl->ready_h[rank][i], ready[rank][i]) == cudaSuccess;
                    for (int root = 0; root < world; ++root)
                        mine = mine && cudaIpcGetEventHandle(&ctl->ack_h[rank][root][i], ack[rank][root][i]) == cudaSuccess;
                } else {
                    ctl->ready_e[rank][i] = ready[rank][i];
                    for (int root = 0; root < world; ++root) ctl->ack_e[rank][root][i] = ack[rank][root][i];
                }
            }
        }
        cudaGetLastError();
        if (!all_ok(mine)) return false;
        mine = true;
        for (int r = 0; r < world; ++r) {
            if (r == rank) continue;
            for (int i = 0; i < kPeerEv; ++i) {
                if (ipc) {
                    mine = mine && cudaIpcOpenEventHandle(&ready[r][i], ctl->ready_h[r][i]) == cudaSuccess;
                    // only the acks addressed to me as root matter
                    mine = mine && cudaIpcOpenEventHandle(&ack[r][rank][i], ctl->ack_h[r][rank][i]) == cudaSuccess;
                } else {
                    ready[r][i] = ctl->ready_e[r][i];
                    ack[r][rank][i] = ctl->ack_e[r][rank][i];
                }
            }
        }
        cudaGetLastError();
        events_ok = all_ok(mine);
        return events_ok;
    }

    // ---- memory: the buffers broadcasts may touch (identical calls on every rank) -----------------------
    bool register_memory(void* const* ptrs, const size_t* sizes, int n) {
        mem_ok = false;
        nalloc = 0;
        if (!events_ok || n > kPeerMaxAlloc) return false;
        bool mine = true;
        for (int a = 0; a < n; ++a) {
            base[rank][a] = static_cast<char*>(ptrs[a]);
            bytes[a] = sizes[a];
            if (ipc)
                mine = mine && cudaIpcGetMemHandle(&ctl->mem_h[rank][a], ptrs[a]) == cudaSuccess;
            else
                ctl->base[rank][a] = ptrs[a];
        }
        cudaGetLastError();
        if (!all_ok(mine)) return false;
        mine = true;
        for (int r = 0; r < world; ++r) {
            if (r == rank) continue;
            if (!ipc) {
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, dev, ctl->dev[r]) != cudaSuccess || !can) {
                    mine = false;
                    continue;
                }
                const cudaError_t e = cudaDeviceEnablePeerAccess(ctl->dev[r], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) mine = false;
                cudaGetLastError();
            }
            for (int a = 0; a < n; ++a) {
                if (ipc) {
                    void* p = nullptr;
                    if (cudaIpcOpenMemHandle(&p, ctl->mem_h[r][a], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                        mine = false;
                        p = nullptr;
                    }
                    base[r][a] = static_cast<char*>(p);
                } else {
                    base[r][a] = static_cast<char*>(ctl->base[r][a]);
                }
            }
        }
        cudaGetLastError();
        nalloc = n;
        mem_ok = all_ok(mine);
        if (!mem_ok) close_peers();
        return mem_ok;
    }
    void close_peers() {
        if (ipc)
            for (int r = 0; r < world; ++r)
                if (r != rank)
                    for (int a = 0; a < nalloc; ++a)
                        if (base[r][a]) {
                            cudaIpcCloseMemHandle(base[r][a]);
                            base[r][a] = nullptr;
                        }
        cudaGetLastError();
    }
    // before the owner frees the buffers: everyone is done with them, then everyone has unmapped them
    void unregister_memory() {
        if (!ctl || nalloc == 0) return;
        const bool was = mem_ok;
        mem_ok = false;
        if (was) barrier(10.0);
        close_peers();
        if (was) barrier(10.0);
        nalloc = 0;
    }

    int find_alloc(const void* p, size_t n, size_t* off) const {
        const char* c = static_cast<const char*>(p);
        for (int a = 0; a < nalloc; ++a)
            if (c >= base[rank][a] && c + n <= base[rank][a] + bytes[a]) {
                *off = (size_t)(c - base[rank][a]);
                return a;
            }
        return -1;
    }
    bool usable(const void* p, size_t n, size_t min_bytes) const {
        size_t off;
        return mem_ok && n >= min_bytes && find_alloc(p, n, &off) >= 0;
    }

    // One broadcast of n bytes at p (the same buffer offset on every rank) from `root`, on `stream`.
    // false: the rendezvous timed out or a CUDA call failed (err says which).
    bool bcast(void* p, size_t n, int root, cudaStream_t stream) {
        size_t off = 0;
        const int a = find_alloc(p, n, &off);
        const uint64_t s = nroot[root]++;
        const int slot = (int)(s % kPeerEv);
        if (broken) return false;
        if (debug) fprintf(stderr, "[peer %d] bcast root %d seq %llu alloc %d off %zu bytes %zu\n", rank, root,
                           (unsigned long long)s, a, off, n);
        if (rank == root) {
            if (cudaEventRecord(ready[rank][slot], stream) != cudaSuccess) return fail("cudaEventRecord(ready)");
            ctl->posted[rank].store(s + 1, std::memory_order_release);
            for (int r = 0; r < world; ++r) {
                if (r == rank) continue;
                if (!wait_ge(ctl->acked[r][rank], s + 1, timeout_s)) return fail("a receiver did not arrive (ack)");
                if (cudaStreamWaitEvent(stream, ack[r][rank][slot], 0) != cudaSuccess) return fail("cudaStreamWaitEvent(ack)");
            }
        } else {
            if (!wait_ge(ctl->posted[root], s + 1, timeout_s)) return fail("the root did not arrive (post)");
            if (cudaStreamWaitEvent(stream, ready[root][slot], 0) != cudaSuccess) return fail("cudaStreamWaitEvent(ready)");
            if (cudaMemcpyAsync(p, base[root][a] + off, n, cudaMemcpyDefault, stream) != cudaSuccess)
                return fail("cudaMemcpyAsync(peer)");
            if (cudaEventRecord(ack[rank][root][slot], stream) != cudaSuccess) return fail("cudaEventRecord(ack)");
            ctl->acked[rank][root].store(s + 1, std::memory_order_release);
            pulled_bytes += (int64_t)n;
        }
        return true;
    }
    bool fail(const char* what) {
        broken = true;
        if (debug) fprintf(stderr, "[peer %d] FAILED: %s\n", rank, what);
        err = std::string("peer broadcast: ") + what;
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) err += std::string(": ") + cudaGetErrorString(e);
        return false;
    }

    void detach() {
        unregister_memory();
        for (int i = 0; i < kPeerEv; ++i) {
            if (ready[rank][i]) cudaEventDestroy(ready[rank][i]);
            for (int root = 0; root < world; ++root)
                if (ack[rank][root][i]) cudaEventDestroy(ack[rank][root][i]);
            if (ipc)
                for (int r = 0; r < world; ++r)
                    if (r != rank) {
                        if (ready[r][i]) cudaEventDestroy(ready[r][i]);
                        if (ack[r][rank][i]) cudaEventDestroy(ack[r][rank][i]);
                    }
        }
        cudaGetLastError();
        if (ipc && ctl) munmap(ctl, sizeof(PeerCtl));
        ctl = nullptr;
        events_ok = false;
    }
};

}  // namespace gogp
