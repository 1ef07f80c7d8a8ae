// api.cu — the extern "C" boundary of libb200aqp.so (see include/aqp/b200_aqp.h) and the host-side
// orchestration of the join: context, workspace, pass planning, phase timing, result hand-back in
// the reference's layouts.
//
// Host-side counterpart of join_init_run / prj_thread / RHO
// (Join-Benchmarks/lib/Joins/src/radix/radix_join.cpp:1369-1643, :1067-1356) and of the scan ECALLs
// (Scan-Micro-Benchmarks/microbenchmarks/SimdScanMulti/Enclave/Enclave.cpp:100-133,:270-299).
// The reference's threads, barriers and task queues have no equivalent here: a phase is one kernel
// launch on one stream and stream order is the barrier.
#include <cstddef>
#include <chrono>
#include <cstdarg>
#include <cstring>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "join_internal.cuh"

namespace aqp {

thread_local std::string g_last_error;
unsigned long long g_kernel_launches = 0;
unsigned g_device_epoch = 0;

void set_error(const std::string &msg) {
    g_last_error = msg;
    if (getenv("B200_AQP_DEBUG")) fprintf(stderr, "b200aqp: %s\n", msg.c_str());
}

// registry of every DevBuf (see join_internal.cuh). A function-local static so that it exists before the first
// static DevBuf of any translation unit is constructed.
static std::vector<DevBuf *> &devbuf_list() {
    static std::vector<DevBuf *> *v = new std::vector<DevBuf *>();
    return *v;
}
static std::mutex &devbuf_mu() {
    static std::mutex *m = new std::mutex();
    return *m;
}
void devbuf_register(DevBuf *b, bool add) {
    std::lock_guard<std::mutex> lk(devbuf_mu());
    auto &v = devbuf_list();
    if (add) {
        v.push_back(b);
    } else {
        for (size_t i = 0; i < v.size(); ++i)
            if (v[i] == b) {
                v[i] = v.back();
                v.pop_back();
                break;
            }
    }
}
void devbuf_release_all() {
    std::lock_guard<std::mutex> lk(devbuf_mu());
    for (DevBuf *b : devbuf_list()) b->release();
}

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct Ctx {
    bool inited = false;
    int device = -1;
    int verbose = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};
    DevBuf tmp[4];        // pass-1 R, pass-1 S, pass-2 R, pass-2 S
    DevBuf meta;          // histograms, offsets, cursors, work list, result accumulators
    DevBuf relR, relS;    // H2D copies of host relations (run_join / preload)
    DevBuf out;           // materialised triples
    DevBuf scan_in, scan_out, scan_scratch;
    uint64_t preR = 0, preS = 0;
    bool preloaded = false;
    b200_join_stats_t last = {};
    uint64_t scan_copy_ns = 0;
    double t0 = 0;
};
static Ctx g;
static std::recursive_mutex g_mu;

static int ensure_init() {
    if (g.inited) return 0;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        set_error(std::string("no usable CUDA device (") + cudaGetErrorString(e) +
                  "); libb200aqp has no CPU fallback");
        return -1;
    }
    int dev = 0;
    AQP_CUDA_OK(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    AQP_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) {
        set_error(std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) +
                  std::to_string(prop.minor) + "; libb200aqp is built for sm_100a (B200) only");
        return -1;
    }
    g.device = dev;
    AQP_CUDA_OK(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
    for (auto &e2 : g.ev) AQP_CUDA_OK(cudaEventCreate(&e2));
    g.t0 = now_s();
    g.inited = true;
    return 0;
}

// the reference's logger line format (Join-Benchmarks/lib/Logger/src/Logger.cpp:71-75) without colours
static void log_info(const char *fmt, ...) {
    if (!g.verbose) return;
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    printf("\x1b[32m[%8.4f][ INFO] %s\x1b[0m\n", now_s() - g.t0, buf);   // colour codes as Logger.cpp:71-75: runner.py:52 counts on the trailing reset
}

[[noreturn]] static void die(const char *what) {
    fprintf(stderr, "[b200aqp][ERROR] %s: %s\n", what, g_last_error.c_str());
    exit(EXIT_FAILURE);
}

constexpr uint32_t kSharedCursorMinBits = 7;   // pass-1 fan-out from which shared (atomic) cursors beat CTA-private ones

// ---------------------------------------------------------------------------------------------
// join planning — the GPU analogue of calc_num_radix_bits / calc_num_passes
// (radix_join.cpp:295-329): partitions are sized so the build side of a co-partition fits the
// shared-memory hash table (kBuildCap tuples) instead of a quarter of the CPU's L2.
// ---------------------------------------------------------------------------------------------
// dead_bits: low key bits the caller knows to carry no information (TPC-H order keys use 8 of every 32 values, so
// 2 of the low 5 bits are dead): the digit still is (key & MASK) >> R on raw bits, it just covers that many more
// bits, which keeps the populated partitions at the planned size instead of 2^dead_bits times larger.
static void plan_bits(uint64_t nR, uint32_t *total, uint32_t *b1, uint32_t *b2, uint32_t dead_bits = 0) {
    uint64_t parts = (nR + kBuildCap - 1) / kBuildCap;
    uint32_t bits = 0;
    while ((1ull << bits) < parts) ++bits;
    if (bits) bits += dead_bits;
    if (bits > 2 * kMaxFanoutBits) bits = 2 * kMaxFanoutBits;   // larger build sides use several build rounds
    if (const char *e = getenv("B200_AQP_RADIX_BITS")) {
        int v = atoi(e);
        if (v >= 0 && v <= 2 * kMaxFanoutBits) bits = (uint32_t) v;
    }
    *total = bits;
    if (bits <= (uint32_t) kMaxFanoutBits) {
        *b1 = bits;
        *b2 = 0;
    } else {
        *b1 = bits / 2;   // pass-1 bits = floor(bits / passes), pass 2 takes the rest (:331-337)
        *b2 = bits - *b1;
    }
}

static __global__ void trivial_offsets_kernel(uint32_t *offR, uint32_t nR, uint32_t *offS, uint32_t nS) {
    offR[0] = 0;
    offR[1] = nR;
    offS[0] = 0;
    offS[1] = nS;
}

// chunked_table_t layout on the device (data-types.h:68-92): slot g of the flat result lives at
// chunk g / TUPLES_PER_CHUNK, entry g % TUPLES_PER_CHUNK.
static __global__ void chunkify_kernel(const output_triple_t *flat, unsigned char *chunks, uint64_t n) {
    uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t c = i / TUPLES_PER_CHUNK, k = i % TUPLES_PER_CHUNK;
        output_triple_t t = flat[i];
        uint32_t *dst = reinterpret_cast<uint32_t *>(chunks + c * sizeof(table_chunk_t) + 8 + k * sizeof(output_triple_t));
        dst[0] = t.key;
        dst[1] = t.Rpayload;
        dst[2] = t.Spayload;
        if (k == 0) {
            uint64_t left = n - i;
            *reinterpret_cast<uint64_t *>(chunks + c * sizeof(table_chunk_t)) =
                left < TUPLES_PER_CHUNK ? left : (uint64_t) TUPLES_PER_CHUNK;
        }
    }
}

struct MetaLayout {
    size_t histR, histS, offR, offS, cur1R, cur1S, cur2R, cur2S, segR, segS, tileR, tileS, seg1R, seg1S, item_start,
        items, result, bhR, bhS, bbR, bbS, total, zero_bytes;
};
static MetaLayout meta_layout(uint32_t bits, uint32_t b1, uint64_t nS) {
    const size_t P = (size_t) 1 << bits, F1 = (size_t) 1 << b1;
    MetaLayout m{};
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t r = o;
        o += (bytes + 255) & ~(size_t) 255;
        return r;
    };
    // zeroed region first: histograms + result accumulators
    m.histR = take(P * 4);
    m.histS = take(P * 4);
    m.result = take(sizeof(JoinResult));
    m.zero_bytes = o;
    m.offR = take((P + 1) * 4);
    m.offS = take((P + 1) * 4);
    m.cur1R = take(F1 * 4);
    m.cur1S = take(F1 * 4);
    m.cur2R = take(P * 4);
    m.cur2S = take(P * 4);
    m.segR = take((F1 + 1) * 4);
    m.segS = take((F1 + 1) * 4);
    m.tileR = take((F1 + 1) * 4);
    m.tileS = take((F1 + 1) * 4);
    m.seg1R = take(16);
    m.seg1S = take(16);
    const size_t NB = pass1_blocks();
    m.bhR = take(NB * F1 * 4);
    m.bhS = take(NB * F1 * 4);
    m.bbR = take(NB * F1 * 4);
    m.bbS = take(NB * F1 * 4);
    m.item_start = take((P + 1) * 4);
    m.items = take((nS / kProbeChunk + P + 1) * sizeof(uint2));
    m.total = o;
    return m;
}

// The whole local join on device-resident relations. Phases and their reference counterparts:
//   histogram  -> partition_hist              (radix_join.cpp:617-654)
//   plan       -> prefix sums                 (:886-915)
//   pass 1/2   -> partition_copy / radix_cluster (:659-697, :715-761)
//   join       -> bucket_chaining_join        (:359-458)
// ---------------------------------------------------------------------------------------------
// The histogram-free plan. The reference sizes every partition exactly from a histogram pass (partition_hist,
// radix_join.cpp:617-654) because its partitions are packed back to back; here that pass is 0.8 of the join's 5.6 ms
// (8 of 48 B/tuple). With shared cursors (api.cu: kSharedCursorMinBits) the scatter needs only a START per partition,
// so every partition of both passes gets a REGION of fixed capacity - the mean size plus slack - and the cursors
// run inside the regions: no histogram, no prefix sums. Pass 2 reads the regions of pass 1 as segments with gaps (the
// multi-GPU receive layout, kGapSegment), build/probe takes [begin, end) per partition. A run that does not fit its
// region is dropped and raises a flag; the join is then repeated with exact offsets (join_device_locked below), so any
// input stays correct. To keep skewed inputs from paying for a failed attempt, 1/256 of the lines of both relations is
// histogrammed first (~0.05 ms) and the plan is declined when the sample shows partitions beyond what the regions hold.
// Count/checksum joins from 11 radix bits on (their bits are split 7 + rest for it); B200_AQP_HISTFREE=0 turns it off.
// ---------------------------------------------------------------------------------------------
struct HistFreeLayout {
    size_t result, flag, verdict, sample[2], zero_bytes, cur1[2], seg1[2], seg_off[2], seg_tile[2], cur2[2], beg[2], end[2],
        seg_group, item_start, items, total;
};
static HistFreeLayout histfree_layout(uint32_t bits, uint32_t b1, uint64_t nS) {
    const size_t P = (size_t) 1 << bits, F1 = (size_t) 1 << b1;
    HistFreeLayout m{};
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t r = o;
        o += (bytes + 255) & ~(size_t) 255;
        return r;
    };
    m.result = take(sizeof(JoinResult));
    m.flag = take(16);
    m.verdict = take(32);
    m.sample[0] = take(P * 4);
    m.sample[1] = take(P * 4);
    m.zero_bytes = o;
    for (int r = 0; r < 2; ++r) {
        m.cur1[r] = take(F1 * 4);
        m.seg1[r] = take(16);
        m.seg_off[r] = take((2 * F1 + 1) * 4);
        m.seg_tile[r] = take((2 * F1 + 1) * 4);
        m.cur2[r] = take(P * 4);
        m.beg[r] = take(P * 4);
        m.end[r] = take(P * 4);
    }
    m.seg_group = take(2 * F1 * 4);
    m.item_start = take((P + 1) * 4);
    m.items = take((nS / kProbeChunk + P + 1) * sizeof(uint2));
    m.total = o;
    return m;
}
// Region capacities and the sampled test that goes with them (every kSampleStride-th 128-byte line, fewer lines apart for
// small inputs so that a final partition still sees ~64 samples):
//   pass 1 (128+ partitions, thousands of samples each, sigma <= 2 %): accepted up to 1.12 x the mean, capacity 1.25 x
//   final partitions (~64-128 samples each, sigma ~10 %): the test only looks for heavy hitters - accepted up to 2 x the
//   mean in the sample, capacity 2.5 x. A uniform or dense key stays near 1.0; Zipf 0.5 over 2^27 keys puts +70 % on
//   its hottest partition and still fits; Zipf 1.0 is turned away by the pass-1 test.
// What the sample misses the region limits in the scatter kernel catch (flag -> repeat with exact offsets).
// Pass-1 regions start on 32 KiB boundaries (measured: 1.91 -> 1.87 ms), final ones on 128-byte lines.
static uint64_t region_cap1(uint64_t n, uint64_t parts) {
    const uint64_t mean = (n + parts - 1) / parts;
    return (mean + (mean >> 2) + 256 + 4095) / 4096 * 4096;
}
static uint64_t region_cap2(uint64_t n, uint64_t parts) {
    const uint64_t mean = (n + parts - 1) / parts;
    return (2 * mean + (mean >> 1) + 256 + 15) / 16 * 16;
}

// returns 0 and *overflowed = false when the join is done; *overflowed = true when a region was too small (nothing valid
// in stats then); -1 on errors
static int join_device_histfree(const row_t *dR, uint64_t nR, const row_t *dS, uint64_t nS, b200_join_stats_t *stats,
                                cudaStream_t st, uint32_t bits, uint32_t b1, uint32_t b2, bool *overflowed, bool *declined) {
    *declined = false;
    const uint32_t P = 1u << bits, F1 = 1u << b1;
    const uint64_t n[2] = {nR, nS};
    uint64_t cap1[2], cap2[2];
    for (int r = 0; r < 2; ++r) {
        cap1[r] = region_cap1(n[r], F1);
        cap2[r] = region_cap2(n[r], P);
        if (cap1[r] * F1 >= 0xFFFF0000ull || cap2[r] * P >= 0xFFFF0000ull) {   // 32-bit offsets: exact plan
            *overflowed = true;
            *declined = true;
            return 0;
        }
    }
    const unsigned long long launches0 = g_kernel_launches;
    HistFreeLayout m = histfree_layout(bits, b1, nS);
    if (g.meta.ensure(m.total)) return -1;
    for (int r = 0; r < 2; ++r)
        if (g.tmp[r].ensure(cap1[r] * F1 * sizeof(row_t) + 16) || g.tmp[2 + r].ensure(cap2[r] * P * sizeof(row_t) + 16)) return -1;
    unsigned char *mb = static_cast<unsigned char *>(g.meta.p);
    auto u32 = [&](size_t off) { return reinterpret_cast<uint32_t *>(mb + off); };
    JoinResult *d_res = reinterpret_cast<JoinResult *>(mb + m.result);
    uint32_t *d_flag = u32(m.flag);
    RegionArgs ra{};
    ra.bits1 = b1;
    ra.bits2 = b2;
    ra.seg_group = u32(m.seg_group);
    for (int r = 0; r < 2; ++r)
        ra.rel[r] = RegionRel{(uint32_t) n[r], (uint32_t) cap1[r], (uint32_t) cap2[r], u32(m.cur1[r]), u32(m.seg1[r]),
                              u32(m.seg_off[r]), u32(m.seg_tile[r]), u32(m.cur2[r]), u32(m.beg[r]), u32(m.end[r])};
    const row_t *in[2] = {dR, dS};
    row_t *t1[2] = {static_cast<row_t *>(g.tmp[0].p), static_cast<row_t *>(g.tmp[1].p)};
    row_t *t2[2] = {static_cast<row_t *>(g.tmp[2].p), static_cast<row_t *>(g.tmp[3].p)};

    AQP_CUDA_OK(cudaEventRecord(g.ev[0], st));
    AQP_CUDA_OK(cudaMemsetAsync(mb, 0, m.zero_bytes, st));
    // the sampled test (one small read-back; the plan kernel of the regions is queued behind it meanwhile)
    uint32_t verdict[6] = {};
    uint32_t stride[2];
    for (int r = 0; r < 2; ++r) {
        // >= 64 samples per final partition and >= 2560 per pass-1 partition (sigma 2 % against a 12 % margin), every
        // line at most, every 256th at least
        const uint64_t v2 = n[r] / ((uint64_t) P * 64), v1 = n[r] / ((uint64_t) F1 * 2560), v = v1 < v2 ? v1 : v2;
        stride[r] = (uint32_t) (v < 1 ? 1 : (v > 256 ? 256 : v));
    }
    if (region_sample_device(dR, nR, stride[0], u32(m.sample[0]), dS, nS, stride[1], u32(m.sample[1]), bits, st)) return -1;
    if (region_verdict_device(u32(m.sample[0]), u32(m.sample[1]), b1, b2, u32(m.verdict), st)) return -1;
    AQP_CUDA_OK(cudaMemcpyAsync(verdict, u32(m.verdict), sizeof verdict, cudaMemcpyDeviceToHost, st));
    if (region_init_device(ra, st)) return -1;
    AQP_CUDA_OK(cudaStreamSynchronize(st));
    const bool ignore_sample = getenv("B200_AQP_HISTFREE_NOSAMPLE") != nullptr;   // test hook: reach the overflow path
    for (int r = 0; r < 2 && !ignore_sample; ++r) {
        const uint64_t total = verdict[r * 3], max1 = verdict[r * 3 + 1], max2 = verdict[r * 3 + 2];
        if (max1 * F1 * 100 > total * 112 || max2 * P > 2 * total + (uint64_t) P * 8) {   // (+8 samples: tiny inputs)
            *overflowed = true;
            *declined = true;
            return 0;
        }
    }
    AQP_CUDA_OK(cudaEventRecord(g.ev[1], st));
    for (int r = 0; r < 2; ++r)
        if (radix_scatter_launch(in[r], t1[r], u32(m.seg1[r]), u32(m.seg1[r]) + 2, nullptr, 1, n[r], make_digit(0, b1), b1,
                                 u32(m.cur1[r]), nullptr, 0, 0, st, nullptr, (uint32_t) cap1[r], d_flag))
            return -1;
    AQP_CUDA_OK(cudaEventRecord(g.ev[2], st));
    if (region_plan2_device(ra, st)) return -1;
    for (int r = 0; r < 2; ++r)
        if (radix_scatter_launch(t1[r], t2[r], u32(m.seg_off[r]), u32(m.seg_tile[r]), u32(m.seg_group), 2 * F1, n[r],
                                 make_digit(b1, b2), b2, u32(m.cur2[r]), nullptr, 0, 0, st, nullptr, (uint32_t) cap2[r], d_flag))
            return -1;
    AQP_CUDA_OK(cudaEventRecord(g.ev[3], st));
    if (region_plan3_device(ra, st)) return -1;
    uint2 *d_items = reinterpret_cast<uint2 *>(mb + m.items);
    if (join_items_device(u32(m.beg[0]), u32(m.beg[1]), P, u32(m.item_start), d_items, st, u32(m.end[0]), u32(m.end[1])))
        return -1;
    if (build_probe_device(t2[0], u32(m.beg[0]), t2[1], u32(m.beg[1]), u32(m.item_start), d_items, P,
                           nS / kProbeChunk + P + 1, bits, d_res, nullptr, 0, st, u32(m.end[0]), u32(m.end[1])))
        return -1;
    AQP_CUDA_OK(cudaEventRecord(g.ev[4], st));
    JoinResult h{};
    uint32_t h_flag = 0;   // a region overflowed in one of the four scatter launches: nothing below is valid
    AQP_CUDA_OK(cudaMemcpyAsync(&h, d_res, sizeof h, cudaMemcpyDeviceToHost, st));
    AQP_CUDA_OK(cudaMemcpyAsync(&h_flag, d_flag, sizeof h_flag, cudaMemcpyDeviceToHost, st));
    AQP_CUDA_OK(cudaStreamSynchronize(st));
    if (h_flag) {
        *overflowed = true;
        return 0;
    }
    *overflowed = false;
    b200_join_stats_t s{};
    s.matches = (int64_t) h.matches;
    s.checksum = h.checksum;
    s.keysum = h.keysum;
    s.radix_bits = bits;
    s.num_passes = 2;
    s.bits_pass1 = b1;
    s.bits_pass2 = b2;
    s.plan_flags = B200_PLAN_HISTOGRAM_FREE;
    cudaEventElapsedTime(&s.ms_hist, g.ev[0], g.ev[1]);
    cudaEventElapsedTime(&s.ms_pass1, g.ev[1], g.ev[2]);
    cudaEventElapsedTime(&s.ms_pass2, g.ev[2], g.ev[3]);
    cudaEventElapsedTime(&s.ms_join, g.ev[3], g.ev[4]);
    cudaEventElapsedTime(&s.ms_total, g.ev[0], g.ev[4]);
    s.kernel_launches = (uint32_t) (g_kernel_launches - launches0);
    g.last = s;
    if (stats) *stats = s;
    return 0;
}

static int join_device_locked(const row_t *dR, uint64_t nR, const row_t *dS, uint64_t nS, output_triple_t *d_out,
                              uint64_t out_cap, b200_join_stats_t *stats, cudaStream_t st, bool keep_partitions,
                              uint32_t dead_bits = 0) {
    (void) keep_partitions;
    if (ensure_init()) return -1;
    if (nR >= 0xFFFF0000ull || nS >= 0xFFFF0000ull) {
        set_error("join: relations of 2^32 tuples or more are not supported on one GPU");
        return -1;
    }
    const unsigned long long launches0 = g_kernel_launches;
    uint32_t bits, b1, b2;
    plan_bits(nR, &bits, &b1, &b2, dead_bits);
    if (dead_bits && bits > (uint32_t) kMaxSmemHistBits && nS <= nR && !getenv("B200_AQP_RADIX_BITS")) {
        // The extra bits for dead key bits would push the plan past the shared-memory histogram and the CTA-private
        // cursors of pass 1 (TPC-H Q12 at SF100: 150 M orders, 17 -> 16 bits, pass 1 on global atomic cursors at
        // 1.8 TB/s). With a probe side no larger than the build side it is cheaper to stay at 15 bits and take a few
        // build rounds per co-partition: the probe tuples that get re-read are few.
        bits = kMaxSmemHistBits;
        b1 = bits / 2;
        b2 = bits - b1;
    }
    // Count joins from 11 radix bits on are candidates for the histogram-free plan (below), which needs the shared pass-1
    // cursors and therefore 128 pass-1 partitions: the split becomes 7 + (bits - 7) instead of the reference's
    // floor(bits / 2) + rest (2^26 x 2^28: 2.84 -> 2.48 ms, 2^25 x 2^27: 1.44 -> 1.30 ms).
    const bool histfree_candidate = !(getenv("B200_AQP_HISTFREE") && atoi(getenv("B200_AQP_HISTFREE")) == 0) && !d_out &&
                                    !dead_bits && !getenv("B200_AQP_PASS1") && bits >= 11 && bits <= (uint32_t) kMaxSmemHistBits;
    if (histfree_candidate && b1 < kSharedCursorMinBits) {
        b1 = kSharedCursorMinBits;
        b2 = bits - b1;
    }
    const uint32_t P = 1u << bits, F1 = 1u << b1;
    const int passes = bits == 0 ? 0 : (b2 ? 2 : 1);

    uint32_t plan_flags = 0;
    if (histfree_candidate && passes == 2 && b1 >= kSharedCursorMinBits && 2 * F1 <= (uint32_t) kMaxSegs) {
        bool overflowed = false, declined = false;
        if (join_device_histfree(dR, nR, dS, nS, stats, st, bits, b1, b2, &overflowed, &declined)) return -1;
        if (!overflowed) return 0;
        // the sample showed skew, or (rarely) a region was too small after all: exact offsets from here on
        plan_flags = declined ? B200_PLAN_HISTOGRAM_FREE_DECLINED : B200_PLAN_HISTOGRAM_FREE_OVERFLOWED;
    }

    MetaLayout m = meta_layout(bits, b1, nS);
    if (g.meta.ensure(m.total)) return -1;
    if (passes >= 1 && (g.tmp[0].ensure(nR * sizeof(row_t)) || g.tmp[1].ensure(nS * sizeof(row_t)))) return -1;
    if (passes == 2 && (g.tmp[2].ensure(nR * sizeof(row_t)) || g.tmp[3].ensure(nS * sizeof(row_t)))) return -1;
    unsigned char *mb = static_cast<unsigned char *>(g.meta.p);
    auto u32 = [&](size_t off) { return reinterpret_cast<uint32_t *>(mb + off); };
    JoinResult *d_res = reinterpret_cast<JoinResult *>(mb + m.result);

    AQP_CUDA_OK(cudaEventRecord(g.ev[0], st));
    AQP_CUDA_OK(cudaMemsetAsync(mb, 0, m.zero_bytes, st));

    const row_t *finR = dR, *finS = dS;
    if (passes == 0) {
        trivial_offsets_kernel<<<1, 1, 0, st>>>(u32(m.offR), (uint32_t) nR, u32(m.offS), (uint32_t) nS);
        AQP_LAUNCHED();
        AQP_CUDA_OK(cudaEventRecord(g.ev[1], st));
        AQP_CUDA_OK(cudaEventRecord(g.ev[2], st));
        AQP_CUDA_OK(cudaEventRecord(g.ev[3], st));
    } else {
        // pass-1 scatter geometry: NB CTAs, CTA b owns tiles [b*tpb, (b+1)*tpb) of its relation; the
        // histogram kernel runs with the same geometry so it can emit per-CTA pass-1 histogram rows
        const uint32_t NB = pass1_blocks();
        // Pass-1 cursors. CTA-private (every CTA owns a tile range and, from the per-CTA histogram rows, a private slice of
        // every partition: no atomics) or shared (one global atomicAdd per partition and tile, tiles interleaved over the
        // CTAs like pass 2, the next input tile requested early). Shared wins when the fan-out spreads the atomics over
        // enough addresses - 2^27 x 2^29, 128 partitions: pass 1 1.98 -> 1.86 ms, also under Zipf 0.5 / 1.0 - and loses
        // when it does not - 2^24 x 2^26, 32 partitions: 0.23 -> 0.35 ms (profiles/r02_sweep_join_pass1_cursors.txt).
        bool priv = bits <= (uint32_t) kMaxSmemHistBits && (b1 < kSharedCursorMinBits || dead_bits);   // dead bits: few, overfull bins
        if (const char *e = getenv("B200_AQP_PASS1")) {
            if (!strcmp(e, "shared")) priv = false;
            if (!strcmp(e, "private")) priv = bits <= (uint32_t) kMaxSmemHistBits;
        }
        const uint32_t tpbR = (uint32_t) (((nR + kScatterTile - 1) / kScatterTile + NB - 1) / NB);
        const uint32_t tpbS = (uint32_t) (((nS + kScatterTile - 1) / kScatterTile + NB - 1) / NB);
        if (radix_hist_device(dR, nR, make_digit(0, bits), bits, u32(m.histR), priv ? NB : 0,
                              (uint64_t) tpbR * kScatterTile, b1, u32(m.bhR), st))
            return -1;
        if (radix_hist_device(dS, nS, make_digit(0, bits), bits, u32(m.histS), priv ? NB : 0,
                              (uint64_t) tpbS * kScatterTile, b1, u32(m.bhS), st))
            return -1;
        PlanArgs pa{};
        pa.bits1 = b1;
        pa.bits2 = b2;
        pa.nblocks1 = NB;
        pa.rel[0] = RelPlan{u32(m.histR), u32(m.offR), u32(m.cur1R), u32(m.cur2R), u32(m.segR), u32(m.tileR), u32(m.seg1R),
                            priv ? u32(m.bhR) : nullptr, u32(m.bbR)};
        pa.rel[1] = RelPlan{u32(m.histS), u32(m.offS), u32(m.cur1S), u32(m.cur2S), u32(m.segS), u32(m.tileS), u32(m.seg1S),
                            priv ? u32(m.bhS) : nullptr, u32(m.bbS)};
        if (plan_offsets_device(pa, st)) return -1;
        AQP_CUDA_OK(cudaEventRecord(g.ev[1], st));

        row_t *t1R = static_cast<row_t *>(g.tmp[0].p), *t1S = static_cast<row_t *>(g.tmp[1].p);
        if (radix_scatter_launch(dR, t1R, u32(m.seg1R), u32(m.seg1R) + 2, nullptr, 1, nR, make_digit(0, b1), b1,
                                 u32(m.cur1R), priv ? u32(m.bbR) : nullptr, NB, tpbR, st))
            return -1;
        if (radix_scatter_launch(dS, t1S, u32(m.seg1S), u32(m.seg1S) + 2, nullptr, 1, nS, make_digit(0, b1), b1,
                                 u32(m.cur1S), priv ? u32(m.bbS) : nullptr, NB, tpbS, st))
            return -1;
        AQP_CUDA_OK(cudaEventRecord(g.ev[2], st));
        finR = t1R;
        finS = t1S;
        if (passes == 2) {
            row_t *t2R = static_cast<row_t *>(g.tmp[2].p), *t2S = static_cast<row_t *>(g.tmp[3].p);
            if (radix_scatter_launch(t1R, t2R, u32(m.segR), u32(m.tileR), nullptr, F1, nR, make_digit(b1, b2), b2,
                                     u32(m.cur2R), nullptr, 0, 0, st))
                return -1;
            if (radix_scatter_launch(t1S, t2S, u32(m.segS), u32(m.tileS), nullptr, F1, nS, make_digit(b1, b2), b2,
                                     u32(m.cur2S), nullptr, 0, 0, st))
                return -1;
            finR = t2R;
            finS = t2S;
        }
        AQP_CUDA_OK(cudaEventRecord(g.ev[3], st));
    }

    uint2 *d_items = reinterpret_cast<uint2 *>(mb + m.items);
    if (join_items_device(u32(m.offR), u32(m.offS), P, u32(m.item_start), d_items, st)) return -1;
    if (build_probe_device(finR, u32(m.offR), finS, u32(m.offS), u32(m.item_start), d_items, P,
                           nS / kProbeChunk + P + 1, bits, d_res, d_out, out_cap, st))
        return -1;
    AQP_CUDA_OK(cudaEventRecord(g.ev[4], st));

    JoinResult h{};
    AQP_CUDA_OK(cudaMemcpyAsync(&h, d_res, sizeof h, cudaMemcpyDeviceToHost, st));
    AQP_CUDA_OK(cudaStreamSynchronize(st));

    b200_join_stats_t s{};
    s.matches = (int64_t) h.matches;
    s.checksum = h.checksum;
    s.keysum = h.keysum;
    s.radix_bits = bits;
    s.num_passes = (uint32_t) passes;
    s.bits_pass1 = b1;
    s.bits_pass2 = b2;
    s.plan_flags = plan_flags;
    cudaEventElapsedTime(&s.ms_hist, g.ev[0], g.ev[1]);
    cudaEventElapsedTime(&s.ms_pass1, g.ev[1], g.ev[2]);
    cudaEventElapsedTime(&s.ms_pass2, g.ev[2], g.ev[3]);
    cudaEventElapsedTime(&s.ms_join, g.ev[3], g.ev[4]);
    cudaEventElapsedTime(&s.ms_total, g.ev[0], g.ev[4]);
    s.kernel_launches = (uint32_t) (g_kernel_launches - launches0);
    g.last = s;
    if (stats) *stats = s;
    return 0;
}

void join_plan_internal(uint64_t nR, uint32_t dead_bits, uint32_t *total, uint32_t *b1, uint32_t *b2) {
    plan_bits(nR, total, b1, b2, dead_bits);
}

// entry for the library's other translation units (tpch.cu): same lock, same workspace
int join_device_internal(const row_t *dR, uint64_t nR, const row_t *dS, uint64_t nS, output_triple_t *d_out,
                         uint64_t out_cap, b200_join_stats_t *stats, cudaStream_t st, uint32_t dead_bits) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (ensure_init()) return -1;
    return join_device_locked(dR, nR, dS, nS, d_out, out_cap, stats, st ? st : g.stream, false, dead_bits);
}
cudaStream_t library_stream() {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    return ensure_init() ? nullptr : g.stream;
}

// re-run only build/probe on the partitions left in the workspace by the last join (used when the
// materialisation buffer turned out too small)
static int rerun_probe_locked(uint64_t nS, output_triple_t *d_out, uint64_t out_cap, cudaStream_t st,
                              const row_t *dR, const row_t *dS) {
    const uint32_t bits = g.last.radix_bits, b1 = g.last.bits_pass1;
    const uint32_t P = 1u << bits;
    MetaLayout m = meta_layout(bits, b1, nS);
    unsigned char *mb = static_cast<unsigned char *>(g.meta.p);
    auto u32 = [&](size_t off) { return reinterpret_cast<uint32_t *>(mb + off); };
    JoinResult *d_res = reinterpret_cast<JoinResult *>(mb + m.result);
    AQP_CUDA_OK(cudaMemsetAsync(d_res, 0, sizeof(JoinResult), st));
    const row_t *finR = g.last.num_passes == 0 ? dR : static_cast<row_t *>(g.tmp[g.last.num_passes == 2 ? 2 : 0].p);
    const row_t *finS = g.last.num_passes == 0 ? dS : static_cast<row_t *>(g.tmp[g.last.num_passes == 2 ? 3 : 1].p);
    if (build_probe_device(finR, u32(m.offR), finS, u32(m.offS), u32(m.item_start), reinterpret_cast<uint2 *>(mb + m.items),
                           P, nS / kProbeChunk + P + 1, bits, d_res, d_out, out_cap, st))
        return -1;
    AQP_CUDA_OK(cudaStreamSynchronize(st));
    return 0;
}

// chunk-array pointer -> slab that backs all its chunks (see destroy_table). Slabs are pinned host memory when a
// device is present - the materialised result then comes back as ONE DMA at the PCIe rate - and the last slab a
// caller destroyed is kept for the next materialising join: pinning (or first-touching) 0.8 GB costs more than the
// join and the copy together (measured: 75 ms of an 89 ms run_join at 2^24 x 2^26 went into page faults of a fresh
// malloc'd slab).
struct Slab {
    void *p = nullptr;
    size_t bytes = 0;
    bool pinned = false;
};
static std::unordered_map<void *, Slab> g_slabs;
static Slab g_slab_cache;
constexpr size_t kSlabCacheMax = (size_t) 16 << 30;
static void slab_free(Slab &b) {
    if (b.p) {
        if (b.pinned) cudaFreeHost(b.p);
        else free(b.p);
    }
    b = Slab{};
}
static Slab slab_alloc(size_t bytes) {
    if (g_slab_cache.p && g_slab_cache.bytes >= bytes && g_slab_cache.bytes <= 2 * bytes + (1 << 20)) {
        Slab b = g_slab_cache;
        g_slab_cache = Slab{};
        return b;
    }
    slab_free(g_slab_cache);
    Slab b;
    b.bytes = bytes;
    if (cudaHostAlloc(&b.p, bytes, cudaHostAllocDefault) == cudaSuccess) {
        b.pinned = true;
    } else {
        cudaGetLastError();
        b.p = malloc(bytes);
    }
    return b;
}
static void slab_release(Slab b) {   // destroy_table: keep the newest slab for the next join
    if (b.bytes > kSlabCacheMax) {
        slab_free(b);
        return;
    }
    slab_free(g_slab_cache);
    g_slab_cache = b;
}

static void print_reference_timing_lines(uint64_t nR, uint64_t nS) {
    // radix_join.cpp:252-293 — the lines SGXv2Scripts/scripts/helpers/runner.py:19-53 scrapes. "cycles"
    // are device microseconds x 1000 (a nominal 1 GHz counter), so CPMS=1000 recovers microseconds.
    const b200_join_stats_t &s = g.last;
    auto cyc = [](float ms) { return (unsigned long) (ms * 1e6f); };
    uint64_t n = nR + nS;
    uint64_t us = (uint64_t) (s.ms_total * 1000.0f);
    log_info("Total input tuples : %lu", (unsigned long) n);
    log_info("Result tuples : %lu", (unsigned long) s.matches);
    log_info("Total Join Time (cycles)    : %lu", cyc(s.ms_total));
    log_info("Partition Overall (cycles)  : %lu", cyc(s.ms_hist + s.ms_pass1 + s.ms_pass2));
    log_info("Partition Pass One (cycles) : %lu", cyc(s.ms_hist + s.ms_pass1));
    log_info("Partition One Hist (cycles) : %lu", cyc(s.ms_hist));
    log_info("Partition One Copy (cycles) : %lu", cyc(s.ms_pass1));
    log_info("Partition Pass Two (cycles) : %lu", cyc(s.ms_pass2));
    log_info("Partition Two Hist (cycles) : %lu", 0ul);
    log_info("Partition Two Copy (cycles) : %lu", cyc(s.ms_pass2));
    log_info("Build+Join Overall (cycles) : %lu", cyc(s.ms_join));
    log_info("Pure Join Runtime (us) : %lu ", (unsigned long) us);
    log_info("Throughput (M rec/sec) : %.2lf", us ? (double) n / (double) us : 0.0);
    log_info("H2D copy (us) : %lu ", (unsigned long) (s.ms_h2d * 1000.0f));
    log_info("Checksum : %lu", (unsigned long) s.checksum);
}

// A count-only join on HOST relations is a PCIe transfer (97 ms for 5.4 GB) followed by 6 ms of kernels. The probe
// side is therefore cut into chunks: a copier thread moves R and then S chunk by chunk on its own stream while the
// library's stream joins R with every chunk that has arrived - the join is a sum over disjoint parts of S
// (matches, checksum and keysum add up; radix_join.cpp:1232-1250 sums its threads' results the same way). R is
// partitioned again for every chunk (K x 1.1 ms at 2^27), which hides under the copy; what stays exposed after the
// last byte has landed is one join of R with the last chunk instead of the whole join. Materialising joins are not
// chunked (their cost is the result's way back, and their output buffer may have to grow). Measured at 2^27 x 2^29,
// pinned relations: see e2e.variants.pinned_count_overlapped in the bench line.
static uint32_t host_join_chunks(uint64_t nS, bool materialize) {
    if (materialize) return 1;
    // Opt-in (B200_AQP_E2E_CHUNKS=K): the K-fold partitioning of R also shows in the device time the reference's log
    // lines report (run_join's "Throughput" is kernel time, as in the reference), so the default stays one join.
    uint32_t k = 1;
    if (const char *e = getenv("B200_AQP_E2E_CHUNKS")) {
        int v = atoi(e);
        if (v >= 1 && v <= 64) k = (uint32_t) v;
    }
    while (k > 1 && nS / k < 8192) --k;
    return k;
}

static int join_host_chunked(const table_t *R, const table_t *S, uint32_t K, b200_join_stats_t *out, float *ms_h2d, cudaStream_t st) {
    const uint64_t nR = R->num_tuples, nS = S->num_tuples;
    row_t *dR = static_cast<row_t *>(g.relR.p), *dS = static_cast<row_t *>(g.relS.p);
    auto chunk_begin = [&](uint32_t k) { return k >= K ? nS : (nS / K * k) & ~(uint64_t) 4095; };   // tile-aligned cuts
    std::mutex mu;
    std::condition_variable cv;
    int arrived = -1;   // -1: nothing yet, 0: R, k: R and the first k chunks of S; -2: failed
    double t_copy = 0;
    const int device = g.device;
    std::thread copier([&] {
        bool ok = cudaSetDevice(device) == cudaSuccess;
        cudaStream_t cs = nullptr;
        ok = ok && cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) == cudaSuccess;
        const double t0 = now_s();
        auto publish = [&](int v) {
            std::lock_guard<std::mutex> lk(mu);
            arrived = v;
            cv.notify_all();
        };
        ok = ok && copy_h2d_any(dR, R->tuples, nR * sizeof(row_t), cs) == 0 && cudaStreamSynchronize(cs) == cudaSuccess;
        if (ok) publish(0);
        for (uint32_t k = 0; ok && k < K; ++k) {
            const uint64_t b = chunk_begin(k), e = chunk_begin(k + 1);
            ok = copy_h2d_any(dS + b, S->tuples + b, (e - b) * sizeof(row_t), cs) == 0 && cudaStreamSynchronize(cs) == cudaSuccess;
            if (ok) publish((int) k + 1);
        }
        t_copy = now_s() - t0;
        if (cs) cudaStreamDestroy(cs);
        if (!ok) publish(-2);
    });
    b200_join_stats_t sum{};
    int rc = 0;
    for (uint32_t k = 0; k < K && rc == 0; ++k) {
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return arrived == -2 || arrived >= (int) k + 1; });
            if (arrived == -2) {
                rc = -1;
                break;
            }
        }
        const uint64_t b = chunk_begin(k), e = chunk_begin(k + 1);
        b200_join_stats_t s{};
        rc = join_device_locked(dR, nR, dS + b, e - b, nullptr, 0, &s, st, true);
        sum.matches += s.matches;
        sum.checksum += s.checksum;
        sum.keysum += s.keysum;
        sum.ms_hist += s.ms_hist;
        sum.ms_pass1 += s.ms_pass1;
        sum.ms_pass2 += s.ms_pass2;
        sum.ms_join += s.ms_join;
        sum.ms_total += s.ms_total;
        sum.kernel_launches += s.kernel_launches;
        sum.radix_bits = s.radix_bits;
        sum.num_passes = s.num_passes;
        sum.bits_pass1 = s.bits_pass1;
        sum.bits_pass2 = s.bits_pass2;
    }
    copier.join();
    if (rc || arrived == -2) {
        if (arrived == -2) set_error("host join: H2D copy of a relation failed");
        return -1;
    }
    *ms_h2d = (float) (t_copy * 1e3);
    g.last = sum;
    *out = sum;
    return 0;
}

// host-buffer join: H2D, device join, results back in the reference's layout
static int join_host_locked(const table_t *R, const table_t *S, const joinconfig_t *cfg, result_t *res,
                            bool use_preloaded) {
    if (ensure_init()) return -1;
    cudaStream_t st = g.stream;
    uint64_t nR, nS;
    float ms_h2d = 0;
    uint32_t chunks = 1;
    if (use_preloaded) {
        if (!g.preloaded) {
            set_error("b200_join_preload called without b200_preload_relations");
            return -1;
        }
        nR = g.preR;
        nS = g.preS;
    } else {
        nR = R->num_tuples;
        nS = S->num_tuples;
        if (g.relR.ensure(nR * sizeof(row_t) + 16) || g.relS.ensure(nS * sizeof(row_t) + 16)) return -1;
        chunks = host_join_chunks(nS, cfg && cfg->MATERIALIZE);
        if (chunks == 1) {
            double t = now_s();
            // pinned relations (this library's create_relation_* return pinned memory): one DMA each; pageable ones
            // (a caller's malloc): multi-threaded staging through pinned buffers (hostcopy.cpp)
            if (copy_h2d_any(g.relR.p, R->tuples, nR * sizeof(row_t), st) || copy_h2d_any(g.relS.p, S->tuples, nS * sizeof(row_t), st))
                return -1;
            AQP_CUDA_OK(cudaStreamSynchronize(st));
            ms_h2d = (float) ((now_s() - t) * 1e3);
        }
        g.preloaded = false;
    }
    const row_t *dR = static_cast<row_t *>(g.relR.p), *dS = static_cast<row_t *>(g.relS.p);
    const bool mat = cfg && cfg->MATERIALIZE;

    output_triple_t *d_out = nullptr;
    uint64_t cap = 0;
    if (mat) {
        cap = nS ? nS : 1;   // exact for PK-FK joins; grown below if R has duplicate keys
        if (g.out.ensure(cap * sizeof(output_triple_t))) return -1;
        d_out = static_cast<output_triple_t *>(g.out.p);
    }
    b200_join_stats_t s{};
    if (chunks > 1) {
        if (join_host_chunked(R, S, chunks, &s, &ms_h2d, st)) return -1;
    } else if (join_device_locked(dR, nR, dS, nS, d_out, cap, &s, st, true)) {
        return -1;
    }
    if (mat && (uint64_t) s.matches > cap) {
        cap = (uint64_t) s.matches;
        if (g.out.ensure(cap * sizeof(output_triple_t))) return -1;
        d_out = static_cast<output_triple_t *>(g.out.p);
        if (rerun_probe_locked(nS, d_out, cap, st, dR, dS)) return -1;
    }
    g.last.ms_h2d = ms_h2d;

    res->totalresults = s.matches;
    res->nthreads = cfg ? cfg->NTHREADS : 0;
    res->materialized = mat ? 1 : 0;
    res->result_type = 1;
    res->throughput = s.ms_total > 0 ? (double) (nR + nS) / (s.ms_total * 1e3) : 0.0;   // M tuples/s

    // result table (radix_join.cpp:1556 concatenate(): always a chunked_table_t, empty if !MATERIALIZE)
    chunked_table_t *ct = static_cast<chunked_table_t *>(calloc(1, sizeof(chunked_table_t)));
    if (!ct) {
        set_error("out of host memory");
        return -1;
    }
    if (mat) {
        double t = now_s();
        const uint64_t n = (uint64_t) s.matches;
        const uint64_t nreal = (n + TUPLES_PER_CHUNK - 1) / TUPLES_PER_CHUNK;
        // The reference concatenates one chunk list per thread, each with at least one (possibly empty) chunk
        // (radix_join.cpp:1294-1297, ChunkedTable.cpp:52-60,:147-171), and its consumers count on it: thread t of
        // q19FilterJoinResultsChunked starts at chunks[t] unconditionally (Q19Predicates.hpp:147-151). So a materialised
        // result always has at least NTHREADS chunks; the surplus ones are empty.
        const uint64_t nthreads = cfg && cfg->NTHREADS > 0 ? (uint64_t) cfg->NTHREADS : 1;
        const uint64_t nchunks = nreal > nthreads ? nreal : nthreads;
        const size_t bytes = nchunks * sizeof(table_chunk_t), real_bytes = nreal * sizeof(table_chunk_t);
        Slab sl = slab_alloc(bytes);
        unsigned char *slab = static_cast<unsigned char *>(sl.p);
        table_chunk_t **arr = static_cast<table_chunk_t **>(malloc(sizeof(table_chunk_t *) * nchunks));
        if (!slab || !arr) {
            set_error("out of host memory for the materialised result");
            return -1;
        }
        if (n) {
            // lay the chunks out on the device, then one D2H into one host slab
            if (g.tmp[0].ensure(real_bytes)) return -1;   // partitions are dead by now; reuse their space
            chunkify_kernel<<<kNumSMs * 8, 256, 0, st>>>(d_out, static_cast<unsigned char *>(g.tmp[0].p), n);
            AQP_LAUNCHED();
            if (copy_d2h_any(slab, g.tmp[0].p, real_bytes, st)) return -1;   // one DMA when the slab is pinned, staged otherwise
            AQP_CUDA_OK(cudaStreamSynchronize(st));
        }
        for (uint64_t c = 0; c < nchunks; ++c) {
            arr[c] = reinterpret_cast<table_chunk_t *>(slab + c * sizeof(table_chunk_t));
            if (c >= nreal) arr[c]->num_tuples = 0;
        }
        ct->chunks = arr;
        ct->num_chunks = nchunks;
        ct->chunk_capacity = nchunks;
        ct->current_chunk = nchunks - 1;
        ct->num_tuples = n;
        g_slabs[arr] = sl;
        g.last.ms_materialize_host = (float) ((now_s() - t) * 1e3);
    } else {
        ct->chunks = static_cast<table_chunk_t **>(malloc(sizeof(table_chunk_t *)));
        ct->current_chunk = (uint64_t) -1;   // num_chunks - 1 with num_chunks == 0, as concatenate() leaves it
    }
    res->result = ct;

    log_info("Running RHO (B200) with %u passes and %u radix bits", s.num_passes, s.radix_bits);
    if (mat) log_info("Materializing the output");
    print_reference_timing_lines(nR, nS);
    return 0;
}

}  // namespace aqp

using namespace aqp;

// =================================================================================================
// extern "C" surface
// =================================================================================================
extern "C" {

int b200_init(int device) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (device >= 0) {
        cudaError_t e = cudaSetDevice(device);
        if (e != cudaSuccess) {
            set_error(std::string("cudaSetDevice: ") + cudaGetErrorString(e));
            return -1;
        }
        if (g.inited && g.device != device) {
            set_error("b200_init: context already bound to another device; call b200_shutdown first");
            return -1;
        }
    }
    return ensure_init();
}

void b200_shutdown(void) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (!g.inited) return;
    cudaStreamSynchronize(g.stream);
    for (auto &b : g.tmp) b.release();
    g.meta.release();
    g.relR.release();
    g.relS.release();
    g.out.release();
    g.scan_in.release();
    g.scan_out.release();
    g.scan_scratch.release();
    scan_release();
    hostcopy_release();
    slab_free(g_slab_cache);
    devbuf_release_all();   // every workspace of every translation unit, incl. function-local statics
    for (auto &e : g.ev) cudaEventDestroy(e);
    cudaStreamDestroy(g.stream);
    g.stream = nullptr;
    g.inited = false;
    ++g_device_epoch;
    g.preloaded = false;
}

const char *b200_last_error(void) { return g_last_error.c_str(); }
void b200_set_verbose(int level) { g.verbose = level; }
uint64_t b200_kernel_launch_count(void) { return g_kernel_launches; }

void *b200_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (ensure_init()) return nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void b200_host_free(void *p) {
    if (p) cudaFreeHost(p);
}
void *b200_device_alloc(size_t bytes) {
    void *p = nullptr;
    if (ensure_init()) return nullptr;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        set_error(std::string("cudaMalloc: ") + cudaGetErrorString(e));
        return nullptr;
    }
    return p;
}
void b200_device_free(void *p) {
    if (p) cudaFree(p);
}
int b200_memcpy_h2d(void *dst, const void *src, size_t bytes) {
    if (ensure_init()) return -1;
    AQP_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g.stream));
    AQP_CUDA_OK(cudaStreamSynchronize(g.stream));
    return 0;
}
int b200_memcpy_d2h(void *dst, const void *src, size_t bytes) {
    if (ensure_init()) return -1;
    AQP_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g.stream));
    AQP_CUDA_OK(cudaStreamSynchronize(g.stream));
    return 0;
}
int b200_device_sync(void) {
    if (ensure_init()) return -1;
    AQP_CUDA_OK(cudaDeviceSynchronize());
    return 0;
}

// ---- join ---------------------------------------------------------------------------------------
struct result_t *RHO(const struct table_t *relR, const struct table_t *relS, const struct joinconfig_t *config) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    result_t *res = static_cast<result_t *>(calloc(1, sizeof(result_t)));
    if (!res || join_host_locked(relR, relS, config, res, false)) die("RHO");
    return res;
}

void run_join(struct result_t *res, const struct table_t *relR, const struct table_t *relS,
              const char *algorithm_name, const struct joinconfig_t *config) {
    if (!algorithm_name || strcmp(algorithm_name, "RHO") != 0) {
        // joins.cpp:70-73: unknown algorithm -> log + exit
        fprintf(stderr, "[b200aqp][ERROR] Algorithm not found: %s (this library serves RHO only)\n",
                algorithm_name ? algorithm_name : "(null)");
        exit(EXIT_FAILURE);
    }
    result_t *tmp = RHO(relR, relS, config);
    memcpy(res, tmp, sizeof(result_t));   // joins.cpp:74-77
    free(tmp);
}

void destroy_table(struct chunked_table_t *table) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (!table) return;
    auto it = table->chunks ? g_slabs.find(table->chunks) : g_slabs.end();
    if (it != g_slabs.end()) {
        slab_release(it->second);   // all chunks live in one slab
        g_slabs.erase(it);
    } else {
        for (uint64_t i = 0; i < table->num_chunks; ++i) free(table->chunks[i]);   // ChunkedTable.cpp:128-136
    }
    free(table->chunks);
    table->chunks = nullptr;
    table->chunk_capacity = 0;
    table->num_chunks = 0;
    table->current_chunk = 0;
}

void b200_last_join_stats(struct b200_join_stats_t *out) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (out) *out = g.last;
}

int b200_preload_relations(const struct table_t *relR, const struct table_t *relS) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (ensure_init()) return -1;
    const uint64_t nR = relR->num_tuples, nS = relS->num_tuples;
    if (g.relR.ensure(nR * sizeof(row_t) + 16) || g.relS.ensure(nS * sizeof(row_t) + 16)) return -1;
    if (copy_h2d_any(g.relR.p, relR->tuples, nR * sizeof(row_t), g.stream) ||
        copy_h2d_any(g.relS.p, relS->tuples, nS * sizeof(row_t), g.stream))
        return -1;
    AQP_CUDA_OK(cudaStreamSynchronize(g.stream));
    g.preR = nR;
    g.preS = nS;
    g.preloaded = true;
    return 0;
}

int b200_join_preload(const char *algorithm_name, const struct joinconfig_t *config, struct result_t *res) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (!algorithm_name || strcmp(algorithm_name, "RHO") != 0) {
        set_error("b200_join_preload: only RHO is served");
        return -1;
    }
    return join_host_locked(nullptr, nullptr, config, res, true);
}

void b200_free_preload(void) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    g.relR.release();
    g.relS.release();
    g.preloaded = false;
}

int b200_join_device(const struct row_t *d_R, uint64_t nR, const struct row_t *d_S, uint64_t nS,
                     struct output_triple_t *d_out, uint64_t out_capacity, struct b200_join_stats_t *stats,
                     void *stream) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (ensure_init()) return -1;
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : g.stream;
    return join_device_locked(d_R, nR, d_S, nS, d_out, out_capacity, stats, st, false);
}

void b200_join_plan(uint64_t nR, uint32_t *total_bits, uint32_t *bits_pass1, uint32_t *bits_pass2) {
    uint32_t t, a, b;
    plan_bits(nR, &t, &a, &b);
    if (total_bits) *total_bits = t;
    if (bits_pass1) *bits_pass1 = a;
    if (bits_pass2) *bits_pass2 = b;
}

int b200_radix_hist_device(const struct row_t *d_in, uint64_t n, uint32_t shift, uint32_t bits, uint32_t *d_hist,
                           void *stream) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (ensure_init()) return -1;
    return radix_hist_device(d_in, n, make_digit(shift, bits), bits, d_hist, 0, 0, 0, nullptr,
                             stream ? static_cast<cudaStream_t>(stream) : g.stream);
}

int b200_exclusive_scan_u32_device(const uint32_t *d_in, uint32_t n, uint32_t *d_out, void *stream) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (ensure_init()) return -1;
    return exclusive_scan_u32_device(d_in, n, d_out, stream ? static_cast<cudaStream_t>(stream) : g.stream);
}

int b200_radix_scatter_device(const struct row_t *d_in, uint64_t n, uint32_t shift, uint32_t bits,
                              const uint32_t *d_offsets, uint32_t *d_cursors, struct row_t *d_out, void *stream) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (ensure_init()) return -1;
    if (n >= 0xFFFF0000ull || bits > (uint32_t) kMaxFanoutBits) {
        set_error("radix_scatter: n must be < 2^32 and bits <= 8");
        return -1;
    }
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : g.stream;
    // the 4-entry single-segment table lives behind the cursors in the caller's scratch? No: keep it in
    // our own meta-independent buffer so the caller's d_cursors stays exactly 2^bits entries.
    static DevBuf seg;
    if (seg.ensure(64)) return -1;
    uint32_t *d_seg = static_cast<uint32_t *>(seg.p);
    if (single_segment_setup((uint32_t) n, d_offsets, 1u << bits, d_cursors, d_seg, st)) return -1;
    return radix_scatter_launch(d_in, d_out, d_seg, d_seg + 2, nullptr, 1, n, make_digit(shift, bits), bits, d_cursors,
                                nullptr, 0, 0, st);
}

// ---- sharded (multi-GPU) join stages ----------------------------------------------------------------
int b200_shard_pass1_device(const struct row_t *d_in, uint64_t n, uint32_t total_bits, uint32_t bits1,
                            uint32_t log2_gpus, struct row_t *d_send, uint32_t *d_hist, uint32_t *d_part1_off,
                            void *stream) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (ensure_init()) return -1;
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : g.stream;
    if (total_bits < bits1 || bits1 > (uint32_t) kMaxFanoutBits || total_bits > (uint32_t) kMaxSmemHistBits ||
        log2_gpus > bits1 || n >= 0xFFFF0000ull) {
        set_error("b200_shard_pass1_device: need log2_gpus <= bits1 <= 8, bits1 <= total_bits <= 15, n < 2^32");
        return -1;
    }
    const uint32_t NB = pass1_blocks(), F1 = 1u << bits1;
    static DevBuf ws;   // per-CTA histogram rows, private cursors, single-segment table
    const size_t rows = (size_t) NB * F1 * 4;
    if (ws.ensure(2 * rows + 256)) return -1;
    uint32_t *bh = static_cast<uint32_t *>(ws.p), *bb = bh + (size_t) NB * F1, *seg1 = bb + (size_t) NB * F1;
    const uint32_t tpb = (uint32_t) (((n + kScatterTile - 1) / kScatterTile + NB - 1) / NB);
    AQP_CUDA_OK(cudaMemsetAsync(d_hist, 0, sizeof(uint32_t) << total_bits, st));
    if (radix_hist_device(d_in, n, make_digit(0, total_bits, bits1, log2_gpus), total_bits, d_hist, NB,
                          (uint64_t) tpb * kScatterTile, bits1, bh, st))
        return -1;
    if (plan_pass1_device(d_hist, bits1, total_bits - bits1, d_part1_off, seg1, bh, bb, NB, st)) return -1;
    return radix_scatter_launch(d_in, d_send, seg1, seg1 + 2, nullptr, 1, n, make_digit(0, bits1, bits1, log2_gpus),
                                bits1, nullptr, bb, NB, tpb, st);
}

// per-relation workspace of the fused path: per-CTA histogram rows survive from the histogram call to
// the scatter call (slot 0 = R, 1 = S)
struct ShardSlot {
    DevBuf ws;
    uint32_t tpb = 0, bits1 = 0, lg = 0;
    uint64_t n = 0;
};
static ShardSlot g_shard[2];

int b200_shard_hist_device(const struct row_t *d_in, uint64_t n, uint32_t total_bits, uint32_t bits1,
                           uint32_t log2_gpus, uint32_t *d_hist, uint32_t *d_counts1, int slot, void *stream) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (ensure_init()) return -1;
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : g.stream;
    if (slot < 0 || slot > 1 || total_bits < bits1 || bits1 > (uint32_t) kMaxFanoutBits ||
        total_bits > (uint32_t) kMaxSmemHistBits || log2_gpus > bits1 || log2_gpus > 3 || n >= 0xFFFF0000ull) {
        set_error("b200_shard_hist_device: need slot in {0,1}, log2_gpus <= min(bits1,3), bits1 <= 8, total_bits <= 15");
        return -1;
    }
    ShardSlot &sl = g_shard[slot];
    const uint32_t NB = pass1_blocks(), F1 = 1u << bits1;
    if (sl.ws.ensure(2 * (size_t) NB * F1 * 4 + 256)) return -1;
    uint32_t *bh = static_cast<uint32_t *>(sl.ws.p);
    sl.tpb = (uint32_t) (((n + kScatterTile - 1) / kScatterTile + NB - 1) / NB);
    sl.bits1 = bits1;
    sl.lg = log2_gpus;
    sl.n = n;
    AQP_CUDA_OK(cudaMemsetAsync(d_hist, 0, sizeof(uint32_t) << total_bits, st));
    if (radix_hist_device(d_in, n, make_digit(0, total_bits, bits1, log2_gpus), total_bits, d_hist, NB,
                          (uint64_t) sl.tpb * kScatterTile, bits1, bh, st))
        return -1;
    return block_base_device(bh, nullptr, F1, NB, nullptr, d_counts1, nullptr, 0, st);
}

int b200_shard_scatter_device(const struct row_t *d_in, uint64_t n, const uint32_t *d_dest_off, void *const *dest_bufs,
                              int slot, void *stream) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (ensure_init()) return -1;
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : g.stream;
    if (slot < 0 || slot > 1 || g_shard[slot].n != n || !g_shard[slot].ws.p) {
        set_error("b200_shard_scatter_device: call b200_shard_hist_device on the same relation and slot first");
        return -1;
    }
    ShardSlot &sl = g_shard[slot];
    const uint32_t NB = pass1_blocks(), F1 = 1u << sl.bits1;
    uint32_t *bh = static_cast<uint32_t *>(sl.ws.p), *bb = bh + (size_t) NB * F1, *seg1 = bb + (size_t) NB * F1;
    if (block_base_device(bh, d_dest_off, F1, NB, bb, nullptr, seg1, (uint32_t) n, st)) return -1;
    PeerTable pt{};
    pt.n = 1u << sl.lg;
    pt.per_shift = sl.bits1 - sl.lg;
    for (uint32_t i = 0; i < pt.n; ++i) pt.base[i] = static_cast<uint2 *>(dest_bufs[i]);
    return radix_scatter_launch(d_in, nullptr, seg1, seg1 + 2, nullptr, 1, n, make_digit(0, sl.bits1, sl.bits1, sl.lg),
                                sl.bits1, nullptr, bb, NB, sl.tpb, st, &pt);
}

int b200_copy_async(void *dst, const void *src, size_t bytes, void *stream) {
    if (ensure_init()) return -1;
    if (bytes == 0) return 0;
    AQP_CUDA_OK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, stream ? static_cast<cudaStream_t>(stream) : g.stream));
    return 0;
}

int b200_ipc_export(void *d_ptr, unsigned char *handle_out) {
    if (ensure_init()) return -1;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    cudaIpcMemHandle_t h;
    AQP_CUDA_OK(cudaIpcGetMemHandle(&h, d_ptr));
    memcpy(handle_out, &h, sizeof h);
    return 0;
}

int b200_ipc_open(const unsigned char *handle, void **d_ptr_out) {
    if (ensure_init()) return -1;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    AQP_CUDA_OK(cudaIpcOpenMemHandle(d_ptr_out, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

int b200_ipc_close(void *d_ptr) {
    if (ensure_init()) return -1;
    AQP_CUDA_OK(cudaIpcCloseMemHandle(d_ptr));
    return 0;
}

int b200_exchange_plan_device(const uint32_t *d_counts_all, uint32_t world, uint32_t rank, uint32_t bits1,
                              uint32_t bits2, const uint32_t *d_hist_global, uint32_t *d_seg_off, uint32_t *d_dest_off,
                              uint32_t *d_hist_slice, uint64_t *d_host_vals, void *stream) {
    if (ensure_init()) return -1;
    if (world == 0 || world > 8 || (world & (world - 1)) || rank >= world || bits1 > (uint32_t) kMaxFanoutBits ||
        (1u << bits1) < world || bits2 > (uint32_t) kMaxFanoutBits) {
        set_error("b200_exchange_plan_device: need world in {1,2,4,8}, rank < world, world <= 2^bits1, bits <= 8 per pass");
        return -1;
    }
    return exchange_plan_device(d_counts_all, world, rank, bits1, bits2, d_hist_global, d_seg_off, d_dest_off, d_hist_slice,
                                reinterpret_cast<unsigned long long *>(d_host_vals),
                                stream ? static_cast<cudaStream_t>(stream) : g.stream);
}

// workspace layout of the last shard join: lets the materialising form repeat the probe alone with a larger output
struct ShardLast {
    size_t o_res, o_offR, o_offS, o_istart, o_items;
    uint32_t P, hash_shift;
    uint64_t nS;
};
static ShardLast g_shard_last{};
static int shard_join_locked(const struct row_t *d_R, uint64_t nR, const uint32_t *d_segoff_R, const struct row_t *d_S,
                             uint64_t nS, const uint32_t *d_segoff_S, const uint32_t *d_seg_group, uint32_t nseg,
                             uint32_t ngroups, uint32_t shift2, uint32_t bits2, const uint32_t *d_hist_R,
                             const uint32_t *d_hist_S, uint32_t hash_shift, struct b200_join_stats_t *stats,
                             uint64_t *d_result3, void *stream, output_triple_t *d_out = nullptr, uint64_t out_cap = 0);

int b200_shard_join_device(const struct row_t *d_R, uint64_t nR, const uint32_t *d_segoff_R, const struct row_t *d_S,
                           uint64_t nS, const uint32_t *d_segoff_S, const uint32_t *d_seg_group, uint32_t nseg,
                           uint32_t ngroups, uint32_t shift2, uint32_t bits2, const uint32_t *d_hist_R,
                           const uint32_t *d_hist_S, uint32_t hash_shift, struct b200_join_stats_t *stats,
                           void *stream) {
    return shard_join_locked(d_R, nR, d_segoff_R, d_S, nS, d_segoff_S, d_seg_group, nseg, ngroups, shift2, bits2, d_hist_R,
                             d_hist_S, hash_shift, stats, nullptr, stream);
}

int b200_shard_join_async_device(const struct row_t *d_R, uint64_t nR, const uint32_t *d_segoff_R,
                                 const struct row_t *d_S, uint64_t nS, const uint32_t *d_segoff_S,
                                 const uint32_t *d_seg_group, uint32_t nseg, uint32_t ngroups, uint32_t shift2,
                                 uint32_t bits2, const uint32_t *d_hist_R, const uint32_t *d_hist_S, uint32_t hash_shift,
                                 uint64_t *d_result3, void *stream) {
    if (!d_result3) {
        set_error("b200_shard_join_async_device: d_result3 must not be null");
        return -1;
    }
    return shard_join_locked(d_R, nR, d_segoff_R, d_S, nS, d_segoff_S, d_seg_group, nseg, ngroups, shift2, bits2, d_hist_R,
                             d_hist_S, hash_shift, nullptr, d_result3, stream);
}

int b200_shard_join_times(struct b200_join_stats_t *stats) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (ensure_init() || !stats) return -1;
    b200_join_stats_t s = g.last;
    AQP_CUDA_OK(cudaEventSynchronize(g.ev[4]));
    cudaEventElapsedTime(&s.ms_pass2, g.ev[0], g.ev[3]);
    cudaEventElapsedTime(&s.ms_join, g.ev[3], g.ev[4]);
    cudaEventElapsedTime(&s.ms_total, g.ev[0], g.ev[4]);
    g.last = s;
    *stats = s;
    return 0;
}

static int shard_join_locked(const struct row_t *d_R, uint64_t nR, const uint32_t *d_segoff_R, const struct row_t *d_S,
                             uint64_t nS, const uint32_t *d_segoff_S, const uint32_t *d_seg_group, uint32_t nseg,
                             uint32_t ngroups, uint32_t shift2, uint32_t bits2, const uint32_t *d_hist_R,
                             const uint32_t *d_hist_S, uint32_t hash_shift, struct b200_join_stats_t *stats,
                             uint64_t *d_result3, void *stream, output_triple_t *d_out, uint64_t out_cap) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (ensure_init()) return -1;
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : g.stream;
    if (nseg == 0 || nseg > (uint32_t) kMaxSegs || bits2 > (uint32_t) kMaxFanoutBits || nR >= 0xFFFF0000ull ||
        nS >= 0xFFFF0000ull) {
        set_error("b200_shard_join_device: need 1 <= nseg <= 264, bits2 <= 8, relations < 2^32 tuples");
        return -1;
    }
    const unsigned long long launches0 = g_kernel_launches;
    const uint32_t P = ngroups << bits2;
    // workspace: [result][part_off R,S][cursor2 R,S][seg_tile_start R,S][item_start][items]
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t r = o;
        o += (bytes + 255) & ~(size_t) 255;
        return r;
    };
    const size_t o_res = take(sizeof(JoinResult)), o_offR = take((P + 1) * 4), o_offS = take((P + 1) * 4),
                 o_curR = take(P * 4), o_curS = take(P * 4), o_tileR = take((nseg + 1) * 4),
                 o_tileS = take((nseg + 1) * 4), o_istart = take((P + 1) * 4),
                 o_items = take((nS / kProbeChunk + P + 1) * sizeof(uint2));
    if (g.meta.ensure(o) || g.tmp[2].ensure(nR * sizeof(row_t) + 16) || g.tmp[3].ensure(nS * sizeof(row_t) + 16)) return -1;
    unsigned char *mb = static_cast<unsigned char *>(g.meta.p);
    auto u32 = [&](size_t off) { return reinterpret_cast<uint32_t *>(mb + off); };
    JoinResult *d_res = reinterpret_cast<JoinResult *>(mb + o_res);
    AQP_CUDA_OK(cudaEventRecord(g.ev[0], st));
    AQP_CUDA_OK(cudaMemsetAsync(d_res, 0, sizeof(JoinResult), st));
    ShardPlanArgs pa{};
    pa.nparts = P;
    pa.nseg = nseg;
    pa.seg_group = d_seg_group;
    pa.rel[0] = ShardRelPlan{d_hist_R, d_segoff_R, u32(o_offR), u32(o_curR), u32(o_tileR)};
    pa.rel[1] = ShardRelPlan{d_hist_S, d_segoff_S, u32(o_offS), u32(o_curS), u32(o_tileS)};
    if (plan_shard_device(pa, st)) return -1;
    row_t *t2R = static_cast<row_t *>(g.tmp[2].p), *t2S = static_cast<row_t *>(g.tmp[3].p);
    if (radix_scatter_launch(d_R, t2R, d_segoff_R, u32(o_tileR), d_seg_group, nseg, nR, make_digit(shift2, bits2), bits2,
                             u32(o_curR), nullptr, 0, 0, st))
        return -1;
    if (radix_scatter_launch(d_S, t2S, d_segoff_S, u32(o_tileS), d_seg_group, nseg, nS, make_digit(shift2, bits2), bits2,
                             u32(o_curS), nullptr, 0, 0, st))
        return -1;
    AQP_CUDA_OK(cudaEventRecord(g.ev[3], st));
    uint2 *d_items = reinterpret_cast<uint2 *>(mb + o_items);
    if (join_items_device(u32(o_offR), u32(o_offS), P, u32(o_istart), d_items, st)) return -1;
    if (build_probe_device(t2R, u32(o_offR), t2S, u32(o_offS), u32(o_istart), d_items, P, nS / kProbeChunk + P + 1,
                           hash_shift, d_res, d_out, out_cap, st))
        return -1;
    AQP_CUDA_OK(cudaEventRecord(g.ev[4], st));
    g_shard_last = ShardLast{o_res, o_offR, o_offS, o_istart, o_items, P, hash_shift, nS};
    if (d_result3) {
        // asynchronous form: {matches, checksum, keysum} stay on the device (the caller all-reduces them there);
        // nothing here waits for the GPU. Phase times: b200_shard_join_times() after the caller's own sync.
        static_assert(offsetof(JoinResult, matches) == 0 && offsetof(JoinResult, keysum) == 16, "JoinResult layout");
        // (the materialising form also hands out word 3: the triples this GPU produced)
        AQP_CUDA_OK(cudaMemcpyAsync(d_result3, d_res, (d_out ? 4 : 3) * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
        b200_join_stats_t s{};
        s.radix_bits = hash_shift;
        s.num_passes = 2;
        s.bits_pass2 = bits2;
        s.kernel_launches = (uint32_t) (g_kernel_launches - launches0);
        g.last = s;
        return 0;
    }
    JoinResult h{};
    AQP_CUDA_OK(cudaMemcpyAsync(&h, d_res, sizeof h, cudaMemcpyDeviceToHost, st));
    AQP_CUDA_OK(cudaStreamSynchronize(st));
    b200_join_stats_t s{};
    s.matches = (int64_t) h.matches;
    s.checksum = h.checksum;
    s.keysum = h.keysum;
    s.radix_bits = hash_shift;
    s.num_passes = 2;
    s.bits_pass2 = bits2;
    cudaEventElapsedTime(&s.ms_pass2, g.ev[0], g.ev[3]);
    cudaEventElapsedTime(&s.ms_join, g.ev[3], g.ev[4]);
    cudaEventElapsedTime(&s.ms_total, g.ev[0], g.ev[4]);
    s.kernel_launches = (uint32_t) (g_kernel_launches - launches0);
    g.last = s;
    if (stats) *stats = s;
    return 0;
}

// ---- generators (device) --------------------------------------------------------------------------
int b200_gen_pk_device(struct row_t *d_rel, uint64_t n_total, uint64_t row_begin, uint64_t n, uint64_t seed,
                       void *stream) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (ensure_init()) return -1;
    return gen_pk_device(d_rel, n_total, row_begin, n, seed, stream ? static_cast<cudaStream_t>(stream) : g.stream);
}
int b200_gen_fk_device(struct row_t *d_rel, uint64_t n_total, uint64_t maxid, uint64_t row_begin, uint64_t n,
                       uint64_t seed, void *stream) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (ensure_init()) return -1;
    return gen_fk_device(d_rel, n_total, maxid, row_begin, n, seed, stream ? static_cast<cudaStream_t>(stream) : g.stream);
}
int b200_gen_zipf_device(struct row_t *d_rel, uint64_t maxid, double zipf_param, uint64_t row_begin, uint64_t n,
                         uint64_t seed, void *stream) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (ensure_init()) return -1;
    return gen_zipf_device(d_rel, maxid, zipf_param, row_begin, n, seed,
                           stream ? static_cast<cudaStream_t>(stream) : g.stream);
}
int b200_set_rowid_payload_device(struct row_t *d_rel, uint64_t row_begin, uint64_t n, void *stream) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (ensure_init()) return -1;
    return set_rowid_payload_device(d_rel, row_begin, n, stream ? static_cast<cudaStream_t>(stream) : g.stream);
}

// ---- scans ----------------------------------------------------------------------------------------
int b200_bitvector_scan_device(uint8_t lo, uint8_t hi, const uint8_t *d_data, size_t n, uint64_t *d_out, void *stream) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (ensure_init()) return -1;
    return bitvector_scan_device(lo, hi, d_data, n, d_out, stream ? static_cast<cudaStream_t>(stream) : g.stream);
}

int b200_scan_count_device(uint8_t lo, uint8_t hi, const uint8_t *d_data, size_t n, uint64_t *d_count, void *stream) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (ensure_init()) return -1;
    return scan_count_device(lo, hi, d_data, n, d_count, stream ? static_cast<cudaStream_t>(stream) : g.stream);
}

int b200_index_scan_device(uint8_t lo, uint8_t hi, const uint8_t *d_data, size_t n, uint64_t id_base,
                           uint64_t *d_out_ids, uint64_t out_capacity, uint64_t *d_count, void *stream) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (ensure_init()) return -1;
    if (g.scan_scratch.ensure(index_scan_scratch_bytes(n))) return -1;
    return index_scan_device(lo, hi, d_data, n, id_base, d_out_ids, out_capacity, d_count, g.scan_scratch.p,
                             stream ? static_cast<cudaStream_t>(stream) : g.stream);
}

int b200_fill_tiled_column_device(uint8_t *d_data, size_t n, uint64_t pos_begin, void *stream) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (ensure_init()) return -1;
    return fill_tiled_column_device(d_data, n, pos_begin, stream ? static_cast<cudaStream_t>(stream) : g.stream);
}

int b200_fill_skewed_column_device(uint8_t *d_data, size_t n, uint64_t pos_begin, uint32_t p_zero_ppm, uint64_t seed,
                                   void *stream) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (ensure_init()) return -1;
    return fill_skewed_column_device(d_data, n, pos_begin, p_zero_ppm, seed,
                                     stream ? static_cast<cudaStream_t>(stream) : g.stream);
}

uint64_t b200_scan_last_copy_ns(void) { return g.scan_copy_ns; }

// Enclave.cpp:270-299 shape: warm-ups, then num_runs timed passes accumulated into *time_cntr
void b200_bitvector_scan_user(uint8_t lo, uint8_t hi, const uint8_t *data, size_t n, uint64_t *output_buffer,
                              uint64_t *time_cntr, size_t num_runs, size_t warmup_runs, int unique_data) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (ensure_init()) die("b200_bitvector_scan_user");
    if (unique_data) {
        num_runs = 1;
        warmup_runs = 0;
    }
    const size_t nblk = n / 64;
    cudaStream_t st = g.stream;
    if (g.scan_in.ensure(n + 64) || g.scan_out.ensure(nblk * 8 + 64)) die("b200_bitvector_scan_user");
    double t = now_s();
    if (copy_h2d_any(g.scan_in.p, data, nblk * 64, st)) die("b200_bitvector_scan_user");
    cudaStreamSynchronize(st);
    uint64_t copy_ns = (uint64_t) ((now_s() - t) * 1e9);
    const uint8_t *d_in = static_cast<const uint8_t *>(g.scan_in.p);
    uint64_t *d_out = static_cast<uint64_t *>(g.scan_out.p);
    for (size_t i = 0; i < warmup_runs; ++i)
        if (bitvector_scan_device(lo, hi, d_in, n, d_out, st)) die("b200_bitvector_scan_user");
    cudaEventRecord(g.ev[5], st);
    for (size_t i = 0; i < num_runs; ++i)
        if (bitvector_scan_device(lo, hi, d_in, n, d_out, st)) die("b200_bitvector_scan_user");
    cudaEventRecord(g.ev[6], st);
    if (cudaStreamSynchronize(st) != cudaSuccess) {
        set_error("bitvector scan kernel failed");
        die("b200_bitvector_scan_user");
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, g.ev[5], g.ev[6]);
    if (time_cntr) *time_cntr += (uint64_t) ((double) ms * 1e6);
    t = now_s();
    if (copy_d2h_any(output_buffer, d_out, nblk * 8, st)) die("b200_bitvector_scan_user");
    if (cudaStreamSynchronize(st) != cudaSuccess) {
        set_error("bitvector scan D2H failed");
        die("b200_bitvector_scan_user");
    }
    g.scan_copy_ns = copy_ns + (uint64_t) ((now_s() - t) * 1e9);
}

// Enclave.cpp:100-133 shape
// ---- the remaining SIMD512 variants on the 8-bit column: sum, value list, dictionary scan --------------------
int b200_scan_sum_device(uint8_t lo, uint8_t hi, const uint8_t *d_data, size_t n, uint64_t *d_sum, void *stream) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (ensure_init()) return -1;
    return scan_sum_device(lo, hi, d_data, n, d_sum, stream ? static_cast<cudaStream_t>(stream) : g.stream);
}

int b200_value_scan_device(uint8_t lo, uint8_t hi, const uint8_t *d_data, size_t n, uint32_t *d_out,
                           uint64_t out_capacity, uint64_t *d_count, void *stream) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (ensure_init()) return -1;
    if (g.scan_scratch.ensure(index_scan_scratch_bytes(n))) return -1;
    return value_scan_device(lo, hi, d_data, n, d_out, out_capacity, d_count, g.scan_scratch.p,
                             stream ? static_cast<cudaStream_t>(stream) : g.stream);
}

// code range of a value predicate over the sorted dictionary, derived exactly like SIMD512.cpp:297-305 (std::find_if
// twice, then narrowed to uint8 — including the wrap-around for predicates outside the dictionary's range)
static void dict_code_range(int64_t predicate_low, int64_t predicate_high, const int64_t *dict, uint8_t *lo, uint8_t *hi) {
    long l = 0;
    while (l < 256 && !(dict[l] >= predicate_low)) ++l;
    long h = l;
    while (h < 256 && !(dict[h] > predicate_high)) ++h;
    *lo = (uint8_t) l;
    *hi = (uint8_t) ((h - 1) & 0xff);
}

int b200_dict_scan_8bit_64bit_device(int64_t predicate_low, int64_t predicate_high, const int64_t *dict,
                                     const uint8_t *d_data, size_t n, int64_t *d_out, uint64_t out_capacity,
                                     uint64_t *d_count, void *stream) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (ensure_init()) return -1;
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : g.stream;
    static DevBuf d_dict;
    if (d_dict.ensure(256 * sizeof(int64_t)) || g.scan_scratch.ensure(index_scan_scratch_bytes(n))) return -1;
    uint8_t lo, hi;
    dict_code_range(predicate_low, predicate_high, dict, &lo, &hi);
    // the dictionary (2 KiB, host memory) travels with the call; the copy is synchronous with respect to the
    // host buffer, so the caller may reuse it right away
    AQP_CUDA_OK(cudaMemcpyAsync(d_dict.p, dict, 256 * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    AQP_CUDA_OK(cudaStreamSynchronize(st));
    return dict_scan_device(lo, hi, static_cast<const int64_t *>(d_dict.p), d_data, n, d_out, out_capacity, d_count,
                            g.scan_scratch.p, st);
}

// host-buffer forms with the reference functions' argument order (SIMD512.hpp:39-84): H2D, kernels, D2H
static int scan_host_common(const char *who, const uint8_t *data, size_t n, const uint8_t **d_in, uint64_t **d_count) {
    if (ensure_init()) return -1;
    static DevBuf cnt;
    if (g.scan_in.ensure(n + 64) || cnt.ensure(16)) return -1;
    if (copy_h2d_any(g.scan_in.p, data, n / 64 * 64, g.stream)) return -1;
    *d_in = static_cast<const uint8_t *>(g.scan_in.p);
    *d_count = static_cast<uint64_t *>(cnt.p);
    (void) who;
    return 0;
}

uint64_t b200_sum(uint8_t lo, uint8_t hi, const uint8_t *data, size_t n) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    const uint8_t *d_in;
    uint64_t *d_cnt, h = 0;
    if (scan_host_common("b200_sum", data, n, &d_in, &d_cnt) || scan_sum_device(lo, hi, d_in, n, d_cnt, g.stream) ||
        cudaMemcpyAsync(&h, d_cnt, 8, cudaMemcpyDeviceToHost, g.stream) != cudaSuccess ||
        cudaStreamSynchronize(g.stream) != cudaSuccess)
        die("b200_sum");
    return h;
}

uint64_t b200_scan(uint8_t lo, uint8_t hi, const uint8_t *data, size_t n, uint32_t *output_buffer, size_t output_capacity) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    const uint8_t *d_in;
    uint64_t *d_cnt, h = 0;
    if (scan_host_common("b200_scan", data, n, &d_in, &d_cnt) || g.scan_out.ensure(output_capacity * 4 + 64) ||
        g.scan_scratch.ensure(index_scan_scratch_bytes(n)) ||
        value_scan_device(lo, hi, d_in, n, static_cast<uint32_t *>(g.scan_out.p), output_capacity, d_cnt,
                          g.scan_scratch.p, g.stream) ||
        cudaMemcpyAsync(&h, d_cnt, 8, cudaMemcpyDeviceToHost, g.stream) != cudaSuccess ||
        cudaStreamSynchronize(g.stream) != cudaSuccess)
        die("b200_scan");
    const uint64_t c = h < output_capacity ? h : output_capacity;
    if (c && (copy_d2h_any(output_buffer, g.scan_out.p, c * 4, g.stream) || cudaStreamSynchronize(g.stream) != cudaSuccess)) die("b200_scan");
    return h;
}

uint64_t b200_dict_scan_8bit_64bit(int64_t predicate_low, int64_t predicate_high, const int64_t *dict, const uint8_t *data,
                                   size_t n, int64_t *output_buffer, size_t output_capacity) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    const uint8_t *d_in;
    uint64_t *d_cnt, h = 0;
    if (scan_host_common("b200_dict_scan_8bit_64bit", data, n, &d_in, &d_cnt) ||
        g.scan_out.ensure(output_capacity * 8 + 64) ||
        b200_dict_scan_8bit_64bit_device(predicate_low, predicate_high, dict, d_in, n, static_cast<int64_t *>(g.scan_out.p),
                                         output_capacity, d_cnt, g.stream) ||
        cudaMemcpyAsync(&h, d_cnt, 8, cudaMemcpyDeviceToHost, g.stream) != cudaSuccess ||
        cudaStreamSynchronize(g.stream) != cudaSuccess)
        die("b200_dict_scan_8bit_64bit");
    const uint64_t c = h < output_capacity ? h : output_capacity;
    if (c && (copy_d2h_any(output_buffer, g.scan_out.p, c * 8, g.stream) || cudaStreamSynchronize(g.stream) != cudaSuccess))
        die("b200_dict_scan_8bit_64bit");
    return h;
}

// ---- the 16- / 32-bit dictionary scans, the explicit-index scan and the scalar twin of the row-id scan -------------
int b200_explicit_index_scan_device(uint8_t lo, uint8_t hi, const uint64_t *d_index, const uint8_t *d_data, size_t n,
                                    uint64_t *d_out, uint64_t out_capacity, uint64_t *d_count, void *stream) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (ensure_init()) return -1;
    return explicit_index_scan_device(lo, hi, d_index, d_data, n, d_out, out_capacity, d_count,
                                      stream ? static_cast<cudaStream_t>(stream) : g.stream);
}

uint64_t b200_explicit_index_scan(uint8_t lo, uint8_t hi, const uint64_t *index, const uint8_t *data, size_t n,
                                  uint64_t *output_buffer, size_t output_capacity) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    const uint8_t *d_in;
    uint64_t *d_cnt, h = 0;
    static DevBuf d_index;
    const size_t index_entries = (n / 64 + 7) * 8;   // block i reads index registers i .. i+7 (SIMD512.cpp:176)
    if (scan_host_common("b200_explicit_index_scan", data, n, &d_in, &d_cnt) || d_index.ensure(index_entries * 8 + 64) ||
        g.scan_out.ensure(output_capacity * 8 + 64) || copy_h2d_any(d_index.p, index, index_entries * 8, g.stream) ||
        explicit_index_scan_device(lo, hi, static_cast<const uint64_t *>(d_index.p), d_in, n,
                                   static_cast<uint64_t *>(g.scan_out.p), output_capacity, d_cnt, g.stream) ||
        cudaMemcpyAsync(&h, d_cnt, 8, cudaMemcpyDeviceToHost, g.stream) != cudaSuccess ||
        cudaStreamSynchronize(g.stream) != cudaSuccess)
        die("b200_explicit_index_scan");
    const uint64_t c = h < output_capacity ? h : output_capacity;
    if (c && (copy_d2h_any(output_buffer, g.scan_out.p, c * 8, g.stream) || cudaStreamSynchronize(g.stream) != cudaSuccess))
        die("b200_explicit_index_scan");
    return h;
}

uint64_t b200_scalar_index_scan(uint8_t lo, uint8_t hi, const uint8_t *data, size_t n, uint64_t *output_buffer,
                                size_t output_capacity) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (ensure_init()) die("b200_scalar_index_scan");
    static DevBuf cnt;
    uint64_t h = 0;
    if (g.scan_in.ensure(n + 64) || cnt.ensure(16) || g.scan_out.ensure(output_capacity * 8 + 64) ||
        g.scan_scratch.ensure(index_scan_scratch_bytes(n)) || copy_h2d_any(g.scan_in.p, data, n, g.stream) ||
        scalar_index_scan_device(lo, hi, static_cast<const uint8_t *>(g.scan_in.p), n, 0, static_cast<uint64_t *>(g.scan_out.p),
                                 output_capacity, static_cast<uint64_t *>(cnt.p), g.scan_scratch.p, g.stream) ||
        cudaMemcpyAsync(&h, cnt.p, 8, cudaMemcpyDeviceToHost, g.stream) != cudaSuccess ||
        cudaStreamSynchronize(g.stream) != cudaSuccess)
        die("b200_scalar_index_scan");
    const uint64_t c = h < output_capacity ? h : output_capacity;
    if (c && (copy_d2h_any(output_buffer, g.scan_out.p, c * 8, g.stream) || cudaStreamSynchronize(g.stream) != cudaSuccess))
        die("b200_scalar_index_scan");
    return h;
}

// predicate on dictionary VALUES -> range of codes, exactly the reference's two std::find_if calls and its casts
// (SIMD512.cpp:539-547,:585-593): both results go through uint16_t - also for 32-bit codes - so a predicate below the
// first or above the last dictionary entry wraps exactly as it does there
static void wide_code_range(int64_t lo, int64_t hi, const int64_t *dict, size_t dict_size, uint32_t *code_lo, uint32_t *code_hi) {
    const int64_t *low = dict;
    while (low < dict + dict_size && !(*low >= lo)) ++low;
    const int64_t *high = low;
    while (high < dict + dict_size && !(*high > hi)) ++high;
    --high;
    *code_lo = (uint16_t) (low - dict);
    *code_hi = (uint16_t) (high - dict);
}

int b200_dict_scan_wide_device(int code_bits, uint32_t code_lo, uint32_t code_hi, const int64_t *d_dict, const void *d_data,
                               size_t n, int64_t *d_out, uint64_t out_capacity, uint64_t *d_count, void *stream) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (ensure_init()) return -1;
    cudaStream_t st = stream ? static_cast<cudaStream_t>(stream) : g.stream;
    if (code_bits == 16)
        return dict_scan16_device(code_lo, code_hi, d_dict, static_cast<const uint16_t *>(d_data), n, d_out, out_capacity, d_count, st);
    if (code_bits == 32)
        return dict_scan32_device(code_lo, code_hi, d_dict, static_cast<const uint32_t *>(d_data), n, d_out, out_capacity, d_count, st);
    set_error("b200_dict_scan_wide_device: code_bits must be 16 or 32 (8-bit codes: b200_dict_scan_8bit_64bit_device)");
    return -1;
}

static uint64_t dict_scan_wide_host(const char *who, int code_bits, int64_t lo, int64_t hi, const int64_t *dict, size_t dict_size,
                                    const void *data, size_t n, int64_t *output_buffer, size_t output_capacity) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (ensure_init()) die(who);
    static DevBuf d_dict, cnt;
    uint32_t clo, chi;
    wide_code_range(lo, hi, dict, dict_size, &clo, &chi);
    const size_t bytes = n * (size_t) (code_bits / 8);
    uint64_t h = 0;
    if (g.scan_in.ensure(bytes + 64) || cnt.ensure(16) || d_dict.ensure(dict_size * 8 + 64) ||
        g.scan_out.ensure(output_capacity * 8 + 64) || copy_h2d_any(g.scan_in.p, data, bytes, g.stream) ||
        copy_h2d_any(d_dict.p, dict, dict_size * 8, g.stream) ||
        b200_dict_scan_wide_device(code_bits, clo, chi, static_cast<const int64_t *>(d_dict.p), g.scan_in.p, n,
                                   static_cast<int64_t *>(g.scan_out.p), output_capacity, static_cast<uint64_t *>(cnt.p), g.stream) ||
        cudaMemcpyAsync(&h, cnt.p, 8, cudaMemcpyDeviceToHost, g.stream) != cudaSuccess ||
        cudaStreamSynchronize(g.stream) != cudaSuccess)
        die(who);
    const uint64_t c = h < output_capacity ? h : output_capacity;
    if (c && (copy_d2h_any(output_buffer, g.scan_out.p, c * 8, g.stream) || cudaStreamSynchronize(g.stream) != cudaSuccess)) die(who);
    return h;
}

uint64_t b200_dict_scan_16bit_64bit(int64_t predicate_low, int64_t predicate_high, const int64_t *dict, const uint16_t *data,
                                    size_t n, int64_t *output_buffer, size_t output_capacity) {
    return dict_scan_wide_host("b200_dict_scan_16bit_64bit", 16, predicate_low, predicate_high, dict, (size_t) 1 << 16, data, n,
                               output_buffer, output_capacity);
}

uint64_t b200_dict_scan_32bit_64bit(int64_t predicate_low, int64_t predicate_high, const int64_t *dict, size_t dict_size,
                                    const uint32_t *data, size_t n, int64_t *output_buffer, size_t output_capacity) {
    return dict_scan_wide_host("b200_dict_scan_32bit_64bit", 32, predicate_low, predicate_high, dict, dict_size, data, n,
                               output_buffer, output_capacity);
}

void b200_index_scan_user(uint8_t lo, uint8_t hi, const uint8_t *data, size_t n, uint64_t *output_buffer,
                          size_t output_capacity, size_t *output_count, uint64_t *time_cntr, size_t num_runs,
                          size_t warmup_runs, int unique_data) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (ensure_init()) die("b200_index_scan_user");
    if (unique_data) {
        num_runs = 1;
        warmup_runs = 0;
    }
    const size_t nblk = n / 64;
    cudaStream_t st = g.stream;
    if (g.scan_in.ensure(n + 64) || g.scan_scratch.ensure(index_scan_scratch_bytes(n))) die("b200_index_scan_user");
    double t = now_s();
    if (copy_h2d_any(g.scan_in.p, data, nblk * 64, st)) die("b200_index_scan_user");
    cudaStreamSynchronize(st);
    uint64_t copy_ns = (uint64_t) ((now_s() - t) * 1e9);
    const uint8_t *d_in = static_cast<const uint8_t *>(g.scan_in.p);
    // size the device output like pre_alloc_per_thread does with SIMD512::count (ResultAllocators.hpp:8-19), untimed
    static DevBuf cnt;
    if (cnt.ensure(16)) die("b200_index_scan_user");
    uint64_t *d_count = static_cast<uint64_t *>(cnt.p);
    uint64_t h_count = 0;
    if (scan_count_device(lo, hi, d_in, n, d_count, st)) die("b200_index_scan_user");
    cudaMemcpyAsync(&h_count, d_count, 8, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    uint64_t cap = h_count < output_capacity ? h_count : output_capacity;
    if (g.scan_out.ensure(cap * 8 + 64)) die("b200_index_scan_user");
    uint64_t *d_out = static_cast<uint64_t *>(g.scan_out.p);
    for (size_t i = 0; i < warmup_runs; ++i)
        if (index_scan_device(lo, hi, d_in, n, 0, d_out, cap, d_count, g.scan_scratch.p, st)) die("b200_index_scan_user");
    cudaEventRecord(g.ev[5], st);
    for (size_t i = 0; i < num_runs; ++i)
        if (index_scan_device(lo, hi, d_in, n, 0, d_out, cap, d_count, g.scan_scratch.p, st)) die("b200_index_scan_user");
    cudaEventRecord(g.ev[6], st);
    if (cudaStreamSynchronize(st) != cudaSuccess) {
        set_error("index scan kernel failed");
        die("b200_index_scan_user");
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, g.ev[5], g.ev[6]);
    if (time_cntr) *time_cntr += (uint64_t) ((double) ms * 1e6);
    t = now_s();
    cudaMemcpyAsync(&h_count, d_count, 8, cudaMemcpyDeviceToHost, st);
    if (cap && copy_d2h_any(output_buffer, d_out, cap * 8, st)) die("b200_index_scan_user");
    if (cudaStreamSynchronize(st) != cudaSuccess) {
        set_error("index scan D2H failed");
        die("b200_index_scan_user");
    }
    if (output_count) *output_count = (size_t) h_count;
    g.scan_copy_ns = copy_ns + (uint64_t) ((now_s() - t) * 1e9);
}

}  // extern "C"

// ---- materialising form of the shard join, for the multi-GPU host (mg.cu) ---------------------------------------
namespace aqp {

// pass 2 + build/probe over received segments like b200_shard_join_async_device, writing this GPU's matches to d_out
// (at most out_cap triples are stored; all are counted). d_result4 = {matches, checksum, keysum, triples produced}.
int shard_join_materialize_internal(const row_t *d_R, uint64_t nR, const uint32_t *d_segoff_R, const row_t *d_S, uint64_t nS,
                                    const uint32_t *d_segoff_S, const uint32_t *d_seg_group, uint32_t nseg, uint32_t ngroups,
                                    uint32_t shift2, uint32_t bits2, const uint32_t *d_hist_R, const uint32_t *d_hist_S,
                                    uint32_t hash_shift, output_triple_t *d_out, uint64_t out_cap, uint64_t *d_result4,
                                    cudaStream_t st) {
    return shard_join_locked(d_R, nR, d_segoff_R, d_S, nS, d_segoff_S, d_seg_group, nseg, ngroups, shift2, bits2, d_hist_R,
                             d_hist_S, hash_shift, nullptr, d_result4, st, d_out, out_cap);
}

// the probe of the last shard join once more, into a buffer that holds every match (the co-partitions are still in the
// workspace) - the reference materialises whatever the probe produces, so a build side with duplicate keys must not
// lose rows (radix_join.cpp:428-447)
int shard_probe_again_internal(output_triple_t *d_out, uint64_t out_cap, uint64_t *d_result4, cudaStream_t st) {
    std::lock_guard<std::recursive_mutex> lk(g_mu);
    if (ensure_init()) return -1;
    const ShardLast &L = g_shard_last;
    if (!L.P || !g.meta.p) {
        set_error("shard_probe_again: no shard join to repeat");
        return -1;
    }
    unsigned char *mb = static_cast<unsigned char *>(g.meta.p);
    auto u32 = [&](size_t off) { return reinterpret_cast<uint32_t *>(mb + off); };
    JoinResult *d_res = reinterpret_cast<JoinResult *>(mb + L.o_res);
    AQP_CUDA_OK(cudaMemsetAsync(d_res, 0, sizeof(JoinResult), st));
    if (build_probe_device(static_cast<row_t *>(g.tmp[2].p), u32(L.o_offR), static_cast<row_t *>(g.tmp[3].p), u32(L.o_offS),
                           u32(L.o_istart), reinterpret_cast<uint2 *>(mb + L.o_items), L.P, L.nS / kProbeChunk + L.P + 1,
                           L.hash_shift, d_res, d_out, out_cap, st))
        return -1;
    AQP_CUDA_OK(cudaMemcpyAsync(d_result4, d_res, 4 * sizeof(uint64_t), cudaMemcpyDeviceToDevice, st));
    return 0;
}

}  // namespace aqp
