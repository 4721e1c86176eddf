// TEST INFRASTRUCTURE: the product's peer-memory broadcast protocol (gogp_b200/csrc/peer_bcast.hpp, unmodified) with
// the ranks as host threads and the handful of CUDA runtime calls it makes replaced by host stand-ins: an event is a
// counter, a stream executes at once, cudaMemcpyAsync is memcpy on host buffers.  What is left is exactly the
// host-level part of the protocol -- the counters, the barrier, the event-slot reuse, who waits for whom -- which is
// where a rendezvous can go wrong; run under ThreadSanitizer, a receiver that read a buffer the root was already
// allowed to overwrite shows up as a data race.  Never linked into libgogp_b200.so.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace mock {
struct Ev {
    std::atomic<long> recorded{0};
};
inline cudaError_t EventCreateWithFlags(cudaEvent_t* e, unsigned) {
    *e = reinterpret_cast<cudaEvent_t>(new Ev());
    return cudaSuccess;
}
inline cudaError_t EventDestroy(cudaEvent_t e) {
    delete reinterpret_cast<Ev*>(e);
    return cudaSuccess;
}
inline cudaError_t EventRecord(cudaEvent_t e, cudaStream_t) {
    reinterpret_cast<Ev*>(e)->recorded.fetch_add(1, std::memory_order_release);
    return cudaSuccess;
}
// a wait binds to the record that has been called by then: the protocol must have made sure there is one
inline std::atomic<long> waits_without_record{0};
inline cudaError_t StreamWaitEvent(cudaStream_t, cudaEvent_t e, unsigned) {
    if (reinterpret_cast<Ev*>(e)->recorded.load(std::memory_order_acquire) == 0) waits_without_record.fetch_add(1);
    return cudaSuccess;
}
inline cudaError_t MemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind, cudaStream_t) {
    std::memcpy(dst, src, n);
    return cudaSuccess;
}
inline cudaError_t DeviceCanAccessPeer(int* can, int, int) {
    *can = 1;
    return cudaSuccess;
}
inline cudaError_t DeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
inline cudaError_t GetLastError() { return cudaSuccess; }
// the process (CUDA IPC) transport is not exercised here: the thread transport shares the objects themselves
inline cudaError_t IpcFail(...) { return cudaErrorNotSupported; }
}  // namespace mock

#define cudaEventCreateWithFlags mock::EventCreateWithFlags
#define cudaEventDestroy mock::EventDestroy
#define cudaEventRecord mock::EventRecord
#define cudaStreamWaitEvent mock::StreamWaitEvent
#define cudaMemcpyAsync mock::MemcpyAsync
#define cudaDeviceCanAccessPeer mock::DeviceCanAccessPeer
#define cudaDeviceEnablePeerAccess mock::DeviceEnablePeerAccess
#define cudaGetLastError mock::GetLastError
#define cudaIpcGetMemHandle mock::IpcFail
#define cudaIpcOpenMemHandle mock::IpcFail
#define cudaIpcCloseMemHandle mock::IpcFail
#define cudaIpcGetEventHandle mock::IpcFail
#define cudaIpcOpenEventHandle mock::IpcFail
#include "../../gogp_b200/csrc/peer_bcast.hpp"

using namespace gogp;

// `steps` rounds of the sweep's traffic on `world` ranks: every round the diagonal-block staging goes out from a
// rotating owner and the panel from each of `pr` process-row roots, at rotating offsets of two panel generations;
// the root overwrites its buffer as soon as the broadcast returns.  Returns 0, or a code saying what went wrong.
extern "C" int peer_protocol_run(int world, int pr, int steps, long words) {
    PeerCtl* ctl = new PeerCtl();
    std::vector<int> rc(world, 0);
    std::vector<std::thread> th;
    mock::waits_without_record.store(0);
    for (int rank = 0; rank < world; ++rank)
        th.emplace_back([&, rank] {
            PeerLink link;
            link.timeout_s = 20.0;
            if (!link.attach(ctl, nullptr, rank, world, rank) || !link.setup_events()) {
                rc[rank] = 1;
                return;
            }
            std::vector<double> panel0(words * pr), panel1(words * pr), dk(words);
            void* ptrs[3] = {panel0.data(), panel1.data(), dk.data()};
            const size_t sizes[3] = {panel0.size() * 8, panel1.size() * 8, dk.size() * 8};
            if (!link.register_memory(ptrs, sizes, 3)) {
                rc[rank] = 2;
                return;
            }
            auto pattern = [](int step, int root, long i) { return (double)(step * 1000003 + root * 7919) + (double)i; };
            for (int k = 0; k < steps && rc[rank] == 0; ++k) {
                const int owner = k % world;
                if (rank == owner)
                    for (long i = 0; i < words; ++i) dk[i] = pattern(k, owner, i);
                if (!link.bcast(dk.data(), (size_t)words * 8, owner, nullptr)) rc[rank] = 3;
                for (long i = 0; i < words && rc[rank] == 0; i += 97)
                    if (dk[i] != pattern(k, owner, i)) rc[rank] = 4;
                if (rank == owner)
                    for (long i = 0; i < words; ++i) dk[i] = -1.0;  // free to overwrite once the broadcast returned
                std::vector<double>& panel = (k & 1) ? panel1 : panel0;
                const int kc = k % (world / pr);
                for (int rr = 0; rr < pr && rc[rank] == 0; ++rr) {
                    const int root = rr * (world / pr) + kc;
                    double* p = panel.data() + (long)rr * words;
                    if (rank == root)
                        for (long i = 0; i < words; ++i) p[i] = pattern(k, root, i) + 0.5;
                    if (!link.bcast(p, (size_t)words * 8, root, nullptr)) rc[rank] = 5;
                    for (long i = 0; i < words && rc[rank] == 0; i += 89)
                        if (p[i] != pattern(k, root, i) + 0.5) rc[rank] = 6;
                }
            }
            link.detach();
        });
    for (auto& t : th) t.join();
    delete ctl;
    for (int r = 0; r < world; ++r)
        if (rc[r]) return 10 * (r + 1) + rc[r];
    if (mock::waits_without_record.load() != 0) return 7;
    return 0;
}

int main(int argc, char** argv) {
    const int world = argc > 1 ? atoi(argv[1]) : 8, pr = argc > 2 ? atoi(argv[2]) : 4, steps = argc > 3 ? atoi(argv[3]) : 200;
    return peer_protocol_run(world, pr, steps, 4096);
}
