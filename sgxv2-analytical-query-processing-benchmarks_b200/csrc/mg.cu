// mg.cu — the multi-GPU host of the sharded RHO join: one process per GPU of one box, written against the
// extern "C" stage calls of this library, NCCL for the small collectives and CUDA IPC for peer memory.
//
// The reference is one shared-memory process; its inter-thread shuffle is the pass-1 scatter into one shared array
// (Join-Benchmarks/lib/Joins/src/radix/radix_join.cpp:901-926, barriers :1106-1109,:1238). Here the same pass is the
// inter-GPU shuffle, and this file is the analogue of join_init_run (:1369-1638) for G processes:
//
//   main stream   hist(R) hist(S) | dest offsets | scatter(R) scatter(S) -> peers' buffers over NVLink | barrier |
//                 pass 2 (R, S) + build/probe on the received partitions | all-reduce of {matches, checksum, keysum}
//   side stream                   | all-gather(counts)  all-reduce(histograms)  region plan |   (needed before pass 2)
//
// Receive buffers are laid out in REGIONS: one fixed region per source rank, sized for the worst case (the source's
// whole shard), inside which the source packs its partitions of that owner tightly. A source therefore derives every
// destination from its own counts and the scatter starts right after the local histogram; the all-gathered counts
// (segment boundaries for the receiver) and the all-reduced full-width histogram (final partition sizes) travel on
// the side stream while the scatter runs. Round 1 did both collectives and a host read-back in front of the scatter
// (0.17 ms of a 2.0 ms join at 8 GPUs). Nothing reads a size on the host before the final result.
//
// NCCL is loaded with dlopen at b200_mg_init: libb200aqp.so has no link-time dependency on it, and inside a process
// that already loaded a libnccl.so.2 (torch) the same copy is used.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

#include "common.cuh"
#include "join_internal.cuh"

namespace aqp {
namespace {

struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
} nccl;

int load_nccl() {
    if (nccl.h) return 0;
    // Order matters inside a process that also uses another NCCL client (torch bundles its own, newer libnccl.so.2):
    // the dynamic loader keeps ONE library per soname, so whichever copy is loaded first serves everybody. (1) a copy that
    // is already loaded, (2) the path in B200_AQP_NCCL_LIB (the Python binding points it at torch's bundled copy, so a
    // later `import torch` finds the version it was built against), (3) the system library.
    nccl.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!nccl.h)
        if (const char *e = getenv("B200_AQP_NCCL_LIB")) nccl.h = dlopen(e, RTLD_NOW | RTLD_LOCAL);
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names)
        if (!nccl.h) nccl.h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
    if (!nccl.h) {
        set_error(std::string("b200_mg: cannot load NCCL: ") + dlerror());
        return -1;
    }
#define AQP_SYM(field, name)                                                      \
    if (!(*reinterpret_cast<void **>(&nccl.field) = dlsym(nccl.h, name))) {       \
        set_error("b200_mg: NCCL symbol missing: " name);                         \
        return -1;                                                                \
    }
    AQP_SYM(GetUniqueId, "ncclGetUniqueId");
    AQP_SYM(CommInitRank, "ncclCommInitRank");
    AQP_SYM(CommDestroy, "ncclCommDestroy");
    AQP_SYM(AllGather, "ncclAllGather");
    AQP_SYM(AllReduce, "ncclAllReduce");
    AQP_SYM(GetErrorString, "ncclGetErrorString");
#undef AQP_SYM
    return 0;
}

#define AQP_NCCL_OK(expr)                                                                               \
    do {                                                                                                \
        ncclResult_t _r = (expr);                                                                       \
        if (_r != ncclSuccess) {                                                                        \
            set_error(std::string(#expr) + " -> " + nccl.GetErrorString(_r));                           \
            return -1;                                                                                  \
        }                                                                                               \
    } while (0)

struct Mg {
    bool on = false;
    int rank = 0, world = 1;
    uint32_t lg = 0, bits = 0, b1 = 0, b2 = 0, F1 = 0, P = 0, per = 0, nseg = 0;
    uint64_t capR = 0, capS = 0;   // region size in tuples (per source, per owner)
    ncclComm_t comm = nullptr;
    cudaStream_t main = nullptr, side = nullptr;
    cudaEvent_t ev[10] = {};
    cudaEvent_t ev_hist = nullptr, ev_side = nullptr;
    void *recvR = nullptr, *recvS = nullptr;   // this rank's receive buffers (device_alloc, exported)
    void *peerR[8] = {}, *peerS[8] = {};
    DevBuf meta, out;   // out: the materialised triples of this rank's co-partitions
    // carved out of meta
    uint32_t *hist = nullptr, *cnt1 = nullptr, *counts_all = nullptr, *dest = nullptr, *seg = nullptr, *seg_group = nullptr,
             *hsl = nullptr, *flag = nullptr;
    unsigned long long *res3 = nullptr;
} mg;
std::mutex mg_mu;

uint32_t log2u(uint32_t x) {
    uint32_t l = 0;
    while ((1u << l) < x) ++l;
    return l;
}

}  // namespace
}  // namespace aqp

using namespace aqp;

extern "C" {

int b200_mg_unique_id(unsigned char *id_out) {
    if (load_nccl()) return -1;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    AQP_NCCL_OK(nccl.GetUniqueId(&id));
    memcpy(id_out, &id, sizeof id);
    return 0;
}

int b200_mg_finalize(void) {
    std::lock_guard<std::mutex> lk(mg_mu);
    if (!mg.on) return 0;
    cudaStreamSynchronize(mg.main);
    cudaStreamSynchronize(mg.side);
    // peers may still be reading our buffers' mappings: leave together
    nccl.AllReduce(mg.flag, mg.flag, 1, ncclUint32, ncclSum, mg.comm, mg.main);
    cudaStreamSynchronize(mg.main);
    for (int g = 0; g < mg.world; ++g) {
        if (g == mg.rank) continue;
        if (mg.peerR[g]) cudaIpcCloseMemHandle(mg.peerR[g]);
        if (mg.peerS[g]) cudaIpcCloseMemHandle(mg.peerS[g]);
    }
    nccl.AllReduce(mg.flag, mg.flag, 1, ncclUint32, ncclSum, mg.comm, mg.main);   // every mapping is closed
    cudaStreamSynchronize(mg.main);
    if (mg.recvR) cudaFree(mg.recvR);
    if (mg.recvS) cudaFree(mg.recvS);
    nccl.CommDestroy(mg.comm);
    for (auto &e : mg.ev) cudaEventDestroy(e);
    cudaEventDestroy(mg.ev_hist);
    cudaEventDestroy(mg.ev_side);
    cudaStreamDestroy(mg.main);
    cudaStreamDestroy(mg.side);
    mg.meta.release();
    mg.out.release();
    mg = Mg{};
    return 0;
}

int b200_mg_init(int rank, int world, const unsigned char *id128, uint64_t nR_total, uint64_t nS_total) {
    if (world < 1) world = 1;
    // worst-case regions: any source may send its whole shard to one owner (no overflow path, any skew)
    return b200_mg_init_caps(rank, world, id128, nR_total, (nR_total + world - 1) / world, (nS_total + world - 1) / world, 0);
}

int b200_mg_init_caps(int rank, int world, const unsigned char *id128, uint64_t nR_total, uint64_t capR, uint64_t capS,
                      uint32_t dead_bits) {
    std::lock_guard<std::mutex> lk(mg_mu);
    if (mg.on) {
        set_error("b200_mg_init: already initialised; call b200_mg_finalize first");
        return -1;
    }
    if (world < 1 || world > 8 || (world & (world - 1)) || rank < 0 || rank >= world) {
        set_error("b200_mg_init: world must be 1, 2, 4 or 8 and 0 <= rank < world");
        return -1;
    }
    if (b200_init(-1) || load_nccl()) return -1;
    mg.rank = rank;
    mg.world = world;
    mg.lg = log2u((uint32_t) world);
    uint32_t bits, b1, b2;
    join_plan_internal(nR_total, dead_bits, &bits, &b1, &b2);
    if (bits > (uint32_t) kMaxSmemHistBits) {   // the shard histogram is one shared-memory table over all 2^bits partitions;
        bits = kMaxSmemHistBits;                // larger build sides take several build rounds per co-partition instead
        b1 = bits / 2;
        b2 = bits - b1;
    }
    if (b1 < mg.lg) {   // pass 1 needs at least log2(world) bits to route on
        b1 = mg.lg;
        if (bits < b1) bits = b1;
        b2 = bits - b1;
    }
    mg.bits = bits;
    mg.b1 = b1;
    mg.b2 = b2;
    mg.F1 = 1u << b1;
    mg.P = 1u << bits;
    mg.per = mg.F1 / (uint32_t) world;
    mg.nseg = (uint32_t) world * (mg.per + 1);
    if (mg.nseg > (uint32_t) kMaxSegs) {
        set_error("b200_mg_init: too many received segments for one pass-2 launch");
        return -1;
    }
    mg.capR = (capR + 63) & ~(uint64_t) 63;   // regions start on 128-byte lines
    mg.capS = (capS + 63) & ~(uint64_t) 63;
    if (mg.capR * world >= 0xFFFF0000ull || mg.capS * world >= 0xFFFF0000ull) {
        set_error("b200_mg_init: relations of 2^32 tuples or more are not supported");
        return -1;
    }
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    AQP_NCCL_OK(nccl.CommInitRank(&mg.comm, world, id, rank));
    AQP_CUDA_OK(cudaStreamCreateWithFlags(&mg.main, cudaStreamNonBlocking));
    AQP_CUDA_OK(cudaStreamCreateWithFlags(&mg.side, cudaStreamNonBlocking));
    for (auto &e : mg.ev) AQP_CUDA_OK(cudaEventCreate(&e));
    AQP_CUDA_OK(cudaEventCreateWithFlags(&mg.ev_hist, cudaEventDisableTiming));
    AQP_CUDA_OK(cudaEventCreateWithFlags(&mg.ev_side, cudaEventDisableTiming));
    // metadata
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t r = o;
        o += (bytes + 255) & ~(size_t) 255;
        return r;
    };
    const size_t n2 = (size_t) mg.per << b2;
    const size_t o_hist = take(2 * mg.P * 4), o_cnt = take(2 * mg.F1 * 4), o_call = take((size_t) world * 2 * mg.F1 * 4),
                 o_dest = take(2 * mg.F1 * 4), o_seg = take(2 * (mg.nseg + 1) * 4), o_grp = take(mg.nseg * 4),
                 o_hsl = take(2 * n2 * 4), o_flag = take(64), o_res = take(128), o_ipc = take((size_t) world * 128);
    if (mg.meta.ensure(o)) return -1;
    unsigned char *mb = static_cast<unsigned char *>(mg.meta.p);
    AQP_CUDA_OK(cudaMemset(mb, 0, o));
    mg.hist = reinterpret_cast<uint32_t *>(mb + o_hist);
    mg.cnt1 = reinterpret_cast<uint32_t *>(mb + o_cnt);
    mg.counts_all = reinterpret_cast<uint32_t *>(mb + o_call);
    mg.dest = reinterpret_cast<uint32_t *>(mb + o_dest);
    mg.seg = reinterpret_cast<uint32_t *>(mb + o_seg);
    mg.seg_group = reinterpret_cast<uint32_t *>(mb + o_grp);
    mg.hsl = reinterpret_cast<uint32_t *>(mb + o_hsl);
    mg.flag = reinterpret_cast<uint32_t *>(mb + o_flag);
    mg.res3 = reinterpret_cast<unsigned long long *>(mb + o_res);
    // receive buffers + IPC exchange (handles all-gathered through NCCL)
    AQP_CUDA_OK(cudaMalloc(&mg.recvR, mg.capR * world * sizeof(row_t) + 256));
    AQP_CUDA_OK(cudaMalloc(&mg.recvS, mg.capS * world * sizeof(row_t) + 256));
    unsigned char *ipc = mb + o_ipc;
    cudaIpcMemHandle_t hh[2];
    AQP_CUDA_OK(cudaIpcGetMemHandle(&hh[0], mg.recvR));
    AQP_CUDA_OK(cudaIpcGetMemHandle(&hh[1], mg.recvS));
    AQP_CUDA_OK(cudaMemcpy(ipc + (size_t) rank * 128, hh, 128, cudaMemcpyHostToDevice));
    AQP_NCCL_OK(nccl.AllGather(ipc + (size_t) rank * 128, ipc, 128, ncclUint8, mg.comm, mg.main));
    AQP_CUDA_OK(cudaStreamSynchronize(mg.main));
    unsigned char all[8 * 128];
    AQP_CUDA_OK(cudaMemcpy(all, ipc, (size_t) world * 128, cudaMemcpyDeviceToHost));
    for (int g = 0; g < world; ++g) {
        if (g == rank) {
            mg.peerR[g] = mg.recvR;
            mg.peerS[g] = mg.recvS;
            continue;
        }
        cudaIpcMemHandle_t h[2];
        memcpy(h, all + (size_t) g * 128, 128);
        AQP_CUDA_OK(cudaIpcOpenMemHandle(&mg.peerR[g], h[0], cudaIpcMemLazyEnablePeerAccess));
        AQP_CUDA_OK(cudaIpcOpenMemHandle(&mg.peerS[g], h[1], cudaIpcMemLazyEnablePeerAccess));
    }
    mg.on = true;
    return 0;
}

}  // extern "C"

// mat: keep the matches of this rank's co-partitions as triples in mg.out (grown and the probe repeated if they do not fit)
static int mg_join_impl(const struct row_t *d_R, uint64_t nR, const struct row_t *d_S, uint64_t nS, struct b200_mg_result_t *out,
                        bool mat, uint64_t *local_rows) {
    std::lock_guard<std::mutex> lk(mg_mu);
    if (!mg.on) {
        set_error("b200_mg_join: call b200_mg_init first");
        return -1;
    }
    if (nR > mg.capR || nS > mg.capS) {
        set_error("b200_mg_join: this rank's shard is larger than ceil(total / world)");
        return -1;
    }
    const int G = mg.world;
    cudaStream_t st = mg.main;
    const uint32_t F1 = mg.F1, P = mg.P;
    const unsigned long long launches0 = g_kernel_launches;
    AQP_CUDA_OK(cudaEventRecord(mg.ev[0], st));
    // ---- 1. local histograms (per-CTA rows stay in the library for the scatter) -------------------------------
    if (b200_shard_hist_device(d_R, nR, mg.bits, mg.b1, mg.lg, mg.hist, mg.cnt1, 0, st) ||
        b200_shard_hist_device(d_S, nS, mg.bits, mg.b1, mg.lg, mg.hist + P, mg.cnt1 + F1, 1, st))
        return -1;
    AQP_CUDA_OK(cudaEventRecord(mg.ev[1], st));
    AQP_CUDA_OK(cudaEventRecord(mg.ev_hist, st));
    // ---- 2. side stream: what the RECEIVER needs before pass 2 -------------------------------------------------
    AQP_CUDA_OK(cudaStreamWaitEvent(mg.side, mg.ev_hist, 0));
    AQP_NCCL_OK(nccl.AllGather(mg.cnt1, mg.counts_all, 2 * F1, ncclUint32, mg.comm, mg.side));
    AQP_NCCL_OK(nccl.AllReduce(mg.hist, mg.hist, 2 * P, ncclUint32, ncclSum, mg.comm, mg.side));   // in place: re-zeroed by the next hist
    if (region_plan_device(mg.counts_all, (uint32_t) G, (uint32_t) mg.rank, mg.b1, mg.b2, mg.hist, mg.capR, mg.capS, mg.seg,
                           mg.seg_group, mg.hsl, mg.side))
        return -1;
    AQP_CUDA_OK(cudaEventRecord(mg.ev_side, mg.side));
    // ---- 3. main stream: destinations from local counts, fused scatter + exchange -----------------------------
    if (region_dest_device(mg.cnt1, (uint32_t) G, (uint32_t) mg.rank, mg.b1, mg.capR, mg.capS, mg.dest, mg.res3 + 4, st)) return -1;
    if (b200_shard_scatter_device(d_R, nR, mg.dest, mg.peerR, 0, st) ||
        b200_shard_scatter_device(d_S, nS, mg.dest + F1, mg.peerS, 1, st))
        return -1;
    AQP_CUDA_OK(cudaEventRecord(mg.ev[2], st));
    // ---- 4. barrier: every rank's stores have landed (and the side stream's plan is ready) --------------------
    AQP_CUDA_OK(cudaStreamWaitEvent(st, mg.ev_side, 0));
    AQP_NCCL_OK(nccl.AllReduce(mg.flag, mg.flag, 1, ncclUint32, ncclSum, mg.comm, st));
    AQP_CUDA_OK(cudaEventRecord(mg.ev[3], st));
    // ---- 5. local pass 2 + build/probe over the received segments, global result ------------------------------
    const size_t n2 = (size_t) mg.per << mg.b2;
    uint64_t out_cap = 0;
    if (mat) {
        // a unique build key gives at most one triple per probe tuple this rank receives - about nS when the shards
        // are even; anything larger is caught after the sync below
        out_cap = nS + (nS >> 2) + 4096;
        if (mg.out.ensure(out_cap * sizeof(output_triple_t))) return -1;
        out_cap = mg.out.cap / sizeof(output_triple_t);
        if (shard_join_materialize_internal(static_cast<const row_t *>(mg.recvR), mg.capR * G, mg.seg,
                                            static_cast<const row_t *>(mg.recvS), mg.capS * G, mg.seg + (mg.nseg + 1),
                                            mg.seg_group, mg.nseg, mg.per, mg.b1, mg.b2, mg.hsl, mg.hsl + n2, mg.bits,
                                            static_cast<output_triple_t *>(mg.out.p), out_cap,
                                            reinterpret_cast<uint64_t *>(mg.res3), st))
            return -1;
    } else if (b200_shard_join_async_device(static_cast<const row_t *>(mg.recvR), mg.capR * G, mg.seg,
                                            static_cast<const row_t *>(mg.recvS), mg.capS * G, mg.seg + (mg.nseg + 1),
                                            mg.seg_group, mg.nseg, mg.per, mg.b1, mg.b2, mg.hsl, mg.hsl + n2, mg.bits,
                                            reinterpret_cast<uint64_t *>(mg.res3), st))
        return -1;
    AQP_CUDA_OK(cudaEventRecord(mg.ev[4], st));
    AQP_NCCL_OK(nccl.AllReduce(mg.res3, mg.res3, 3, ncclUint64, ncclSum, mg.comm, st));
    AQP_CUDA_OK(cudaEventRecord(mg.ev[5], st));
    unsigned long long h[6] = {};
    AQP_CUDA_OK(cudaMemcpyAsync(h, mg.res3, sizeof h, cudaMemcpyDeviceToHost, st));
    AQP_CUDA_OK(cudaStreamSynchronize(st));   // the one host sync of the join
    if (mat) {
        if (h[3] > out_cap) {   // duplicate build keys or a lopsided key distribution: room for all, probe again
            if (mg.out.ensure(h[3] * sizeof(output_triple_t))) return -1;
            if (shard_probe_again_internal(static_cast<output_triple_t *>(mg.out.p), mg.out.cap / sizeof(output_triple_t),
                                           reinterpret_cast<uint64_t *>(mg.res3 + 8), st))
                return -1;
            AQP_CUDA_OK(cudaStreamSynchronize(st));
        }
        if (local_rows) *local_rows = h[3];
    }
    if (out) {
        memset(out, 0, sizeof *out);
        out->matches = h[0];
        out->checksum = h[1];
        out->keysum = h[2];
        out->radix_bits = mg.bits;
        out->bits_pass1 = mg.b1;
        out->bits_pass2 = mg.b2;
        out->world = (uint32_t) G;
        cudaEventElapsedTime(&out->ms_hist, mg.ev[0], mg.ev[1]);
        cudaEventElapsedTime(&out->ms_scatter, mg.ev[1], mg.ev[2]);
        cudaEventElapsedTime(&out->ms_barrier, mg.ev[2], mg.ev[3]);
        cudaEventElapsedTime(&out->ms_local, mg.ev[3], mg.ev[4]);
        cudaEventElapsedTime(&out->ms_reduce, mg.ev[4], mg.ev[5]);
        cudaEventElapsedTime(&out->ms_total, mg.ev[0], mg.ev[5]);
        b200_join_stats_t s{};
        if (b200_shard_join_times(&s) == 0) {
            out->ms_pass2 = s.ms_pass2;
            out->ms_join = s.ms_join;
        }
        out->kernel_launches = (uint32_t) (g_kernel_launches - launches0);
        out->tuples_sent = nR + nS;
        out->tuples_kept = h[4] + h[5];
    }
    return 0;
}

extern "C" {

// sum of n <= 4 host values over the ranks (collective): what a sharded pipeline needs to turn per-rank counts into its answer
int b200_mg_allreduce_u64(uint64_t *values, int n) {
    std::lock_guard<std::mutex> lk(mg_mu);
    if (!mg.on || n < 1 || n > 4 || !values) {
        set_error("b200_mg_allreduce_u64: needs b200_mg_init and 1 <= n <= 4");
        return -1;
    }
    unsigned long long *d = mg.res3 + 12;
    AQP_CUDA_OK(cudaMemcpyAsync(d, values, (size_t) n * 8, cudaMemcpyHostToDevice, mg.main));
    AQP_NCCL_OK(nccl.AllReduce(d, d, (size_t) n, ncclUint64, ncclSum, mg.comm, mg.main));
    AQP_CUDA_OK(cudaMemcpyAsync(values, d, (size_t) n * 8, cudaMemcpyDeviceToHost, mg.main));
    AQP_CUDA_OK(cudaStreamSynchronize(mg.main));
    return 0;
}

int b200_mg_join(const struct row_t *d_R, uint64_t nR, const struct row_t *d_S, uint64_t nS, struct b200_mg_result_t *out) {
    return mg_join_impl(d_R, nR, d_S, nS, out, false, nullptr);
}

int b200_mg_join_materialize(const struct row_t *d_R, uint64_t nR, const struct row_t *d_S, uint64_t nS,
                             const struct output_triple_t **d_triples, uint64_t *local_rows, struct b200_mg_result_t *out) {
    uint64_t rows = 0;
    if (mg_join_impl(d_R, nR, d_S, nS, out, true, &rows)) return -1;
    if (d_triples) *d_triples = static_cast<const output_triple_t *>(mg.out.p);
    if (local_rows) *local_rows = rows;
    return 0;
}

}  // extern "C"
