// hostcopy.cpp — host <-> device copies for the drop-in (host-buffer) entry points.
//
// The reference's callers hand run_join() / the scan ECALLs plain malloc'd memory (App/TEEBench/native.cpp:62-100,
// Allocator.hpp:95-109). A cudaMemcpy from pageable memory goes through the driver's small internal staging buffer
// on ONE thread (~10-25 GB/s); PCIe Gen5 x16 moves ~55 GB/s. So:
//   * memory that is already pinned (cudaHostAlloc / cudaHostRegister, which includes every relation this library's
//     own create_relation_* return) is copied with one cudaMemcpyAsync;
//   * pageable memory is cut into slices, one per helper thread; every thread memcpy's its slice piecewise into its own
//     pair of pinned staging buffers and queues the DMA of a piece on its own stream while it fills the other buffer
//     (D2H: the mirror image). The host side then runs at the sum of the threads' memcpy rates and the DMA engine
//     never waits for a single memcpy thread.
#include <algorithm>
#include <cstring>
#include <mutex>
#include <thread>
#include <unordered_set>
#include <vector>

#include "common.cuh"
#include "join_internal.cuh"

namespace aqp {

namespace {
constexpr size_t kPiece = 8u << 20;          // staging piece: 8 MiB
constexpr size_t kDirectBelow = 4u << 20;    // small copies: the plain path is as good
struct Lane {
    unsigned char *buf[2] = {nullptr, nullptr};
    cudaStream_t st = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
};
struct Pool {
    std::vector<Lane> lanes;
    int device = -1;
} g_pool;
std::mutex g_pool_mu;

bool is_pinned(const void *p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

int ensure_pool() {
    int dev = 0;
    AQP_CUDA_OK(cudaGetDevice(&dev));
    if (!g_pool.lanes.empty() && g_pool.device == dev) return 0;
    hostcopy_release();
    unsigned hw = std::thread::hardware_concurrency();
    int n = (int) std::min<unsigned>(8u, std::max<unsigned>(2u, hw / 2));
    if (const char *e = getenv("B200_AQP_COPY_THREADS")) n = std::max(1, std::min(32, atoi(e)));
    g_pool.lanes.resize(n);
    for (auto &l : g_pool.lanes) {
        for (int b = 0; b < 2; ++b) {
            AQP_CUDA_OK(cudaHostAlloc(reinterpret_cast<void **>(&l.buf[b]), kPiece, cudaHostAllocDefault));
            AQP_CUDA_OK(cudaEventCreateWithFlags(&l.ev[b], cudaEventDisableTiming));
        }
        AQP_CUDA_OK(cudaStreamCreateWithFlags(&l.st, cudaStreamNonBlocking));
    }
    g_pool.device = dev;
    return 0;
}

// one helper thread's slice: [off, off + len) of the transfer
template <bool kToDevice>
void lane_copy(Lane *l, int device, unsigned char *dev, unsigned char *host, size_t len, cudaError_t *err) {
    cudaSetDevice(device);
    cudaError_t e = cudaSuccess;
    size_t done = 0;
    int b = 0;
    size_t pend_off[2] = {0, 0}, pend_len[2] = {0, 0};   // D2H: pieces whose DMA is queued but not yet copied out
    while (done < len && e == cudaSuccess) {
        const size_t n = std::min(kPiece, len - done);
        if (kToDevice) {
            if ((e = cudaEventSynchronize(l->ev[b])) != cudaSuccess) break;   // the DMA that last read this buffer
            memcpy(l->buf[b], host + done, n);
            if ((e = cudaMemcpyAsync(dev + done, l->buf[b], n, cudaMemcpyHostToDevice, l->st)) != cudaSuccess) break;
            e = cudaEventRecord(l->ev[b], l->st);
        } else {
            if (pend_len[b]) {   // drain what the previous DMA into this buffer brought
                if ((e = cudaEventSynchronize(l->ev[b])) != cudaSuccess) break;
                memcpy(host + pend_off[b], l->buf[b], pend_len[b]);
            }
            if ((e = cudaMemcpyAsync(l->buf[b], dev + done, n, cudaMemcpyDeviceToHost, l->st)) != cudaSuccess) break;
            e = cudaEventRecord(l->ev[b], l->st);
            pend_off[b] = done;
            pend_len[b] = n;
        }
        done += n;
        b ^= 1;
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(l->st);
    if (!kToDevice && e == cudaSuccess)
        for (int k = 0; k < 2; ++k, b ^= 1)   // oldest first
            if (pend_len[b]) memcpy(host + pend_off[b], l->buf[b], pend_len[b]);
    *err = e;
}

template <bool kToDevice>
int staged(void *dev, void *host, size_t bytes, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (ensure_pool()) return -1;
    AQP_CUDA_OK(cudaStreamSynchronize(st));   // whatever produced / still reads the device buffer on the caller's stream
    const int n = (int) g_pool.lanes.size();
    const size_t slice = ((bytes + n - 1) / n + 4095) & ~(size_t) 4095;
    std::vector<std::thread> th;
    std::vector<cudaError_t> err(n, cudaSuccess);
    for (int i = 0; i < n; ++i) {
        const size_t off = (size_t) i * slice;
        if (off >= bytes) break;
        const size_t len = std::min(slice, bytes - off);
        th.emplace_back(lane_copy<kToDevice>, &g_pool.lanes[i], g_pool.device, static_cast<unsigned char *>(dev) + off,
                        static_cast<unsigned char *>(host) + off, len, &err[i]);
    }
    for (auto &t : th) t.join();
    for (cudaError_t e : err)
        if (e != cudaSuccess) {
            set_error(std::string("staged host copy: ") + cudaGetErrorString(e));
            return -1;
        }
    return 0;
}
}  // namespace

void hostcopy_release() {
    for (auto &l : g_pool.lanes) {
        for (int b = 0; b < 2; ++b) {
            if (l.buf[b]) cudaFreeHost(l.buf[b]);
            if (l.ev[b]) cudaEventDestroy(l.ev[b]);
        }
        if (l.st) cudaStreamDestroy(l.st);
    }
    g_pool.lanes.clear();
    g_pool.device = -1;
}

// Both return with the copy COMPLETE for pageable memory and QUEUED on `st` for pinned memory.
int copy_h2d_any(void *dev, const void *host, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return 0;
    if (bytes < kDirectBelow || is_pinned(host)) {
        AQP_CUDA_OK(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, st));
        return 0;
    }
    return staged<true>(dev, const_cast<void *>(host), bytes, st);
}
int copy_d2h_any(void *host, const void *dev, size_t bytes, cudaStream_t st) {
    if (bytes == 0) return 0;
    if (bytes < kDirectBelow || is_pinned(host)) {
        AQP_CUDA_OK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, st));
        return 0;
    }
    return staged<false>(const_cast<void *>(dev), host, bytes, st);
}

// ---- pinned relations (create_relation_* / delete_relation, csrc/host_gen.cpp) ----------------------------------
static std::unordered_set<void *> g_pinned;
static std::mutex g_pinned_mu;
void *host_alloc_prefer_pinned(size_t bytes) {
    void *p = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0 && cudaHostAlloc(&p, bytes, cudaHostAllocDefault) == cudaSuccess) {
        std::lock_guard<std::mutex> lk(g_pinned_mu);
        g_pinned.insert(p);
        return p;
    }
    cudaGetLastError();       // no device: plain memory (the generators are host code and stay usable without a GPU)
    return malloc(bytes);
}
void host_free_any(void *p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(g_pinned_mu);
        auto it = g_pinned.find(p);
        if (it != g_pinned.end()) {
            g_pinned.erase(it);
            cudaFreeHost(p);
            return;
        }
    }
    free(p);
}

}  // namespace aqp
