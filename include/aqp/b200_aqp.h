/*
 * aqp/b200_aqp.h — C ABI of libb200aqp.so, the B200 (sm_100a) implementation of the reference's
 * RHO radix hash join and uint8 column scans.
 *
 * Every entry point is extern "C", takes plain pointers / sizes (no C++ or torch types) and cites
 * the reference interface it replaces (paths relative to /root/reference/).  Three layers:
 *
 *   1. Drop-in operator API, HOST buffers      : run_join, RHO, destroy_table, the generators,
 *                                                 b200_bitvector_scan_user / b200_index_scan_user.
 *   2. ECALL-shaped preload/run split          : b200_preload_relations + b200_join_preload
 *                                                 (H2D once, then time the kernels only).
 *   3. Device-resident API, DEVICE pointers    : b200_*_device — what a multi-GPU host (one process
 *                                                 per GPU) and bench.py drive; caller owns memory.
 *
 * Concurrency: one process drives ONE device through this library, one call at a time - like the reference's
 * run_join, which all its callers invoke sequentially. Host-side entry is serialised by a mutex, but the
 * device-pointer calls return with work still queued and they share ONE workspace (partition buffers, scan state,
 * metadata, phase-timing events): issue them on one stream, or synchronise before switching streams. Per-call state a
 * caller passes in (outputs, counters) is its own. b200_shutdown() followed by b200_init(other device) re-binds the
 * library: every cached buffer is released and every per-device kernel attribute is set again.
 *
 * There is no CPU fallback anywhere: if no CUDA device is usable every call fails loudly
 * (non-zero return, or for the void reference-shaped calls a message on stderr and exit(1), the
 * reference's own error convention — Joins/src/util.cpp:12-19, joins.cpp:70-73).
 */
#ifndef AQP_B200_AQP_H
#define AQP_B200_AQP_H

#include <stddef.h>
#include <stdint.h>

#include "data_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------
 * 0. library / device
 * ---------------------------------------------------------------------------------------------- */
/* 0 on success. device < 0 keeps the current CUDA device. Idempotent. */
int b200_init(int device);
/* release every cached device buffer, stream and event of the calling context */
void b200_shutdown(void);
const char *b200_last_error(void);
/* 0 = silent (default), 1 = print the reference's logger lines scraped by
 * Join-Benchmarks/SGXv2Scripts/scripts/helpers/runner.py:19-53 ("Throughput (M rec/sec) : ...") */
void b200_set_verbose(int level);
/* pinned host memory for callers that want full-speed H2D (plain malloc'd memory also works) */
void *b200_host_alloc(size_t bytes);
void b200_host_free(void *p);
void *b200_device_alloc(size_t bytes);
void b200_device_free(void *p);
int b200_memcpy_h2d(void *dst, const void *src, size_t bytes);
int b200_memcpy_d2h(void *dst, const void *src, size_t bytes);
int b200_device_sync(void);

/* ------------------------------------------------------------------------------------------------
 * 1. join — drop-in operator API (host buffers)
 * ---------------------------------------------------------------------------------------------- */
/* Join-Benchmarks/lib/Joins/include/joins.hpp:4-6 (impl joins.cpp:55-78). Only "RHO" is served by
 * this library; any other name -> error message + exit(EXIT_FAILURE) like joins.cpp:70-73. */
void run_join(struct result_t *res, const struct table_t *relR, const struct table_t *relS,
              const char *algorithm_name, const struct joinconfig_t *config);

/* Join-Benchmarks/lib/Joins/include/radix/radix_join.h:30 (impl radix_join.cpp:1640-1643).
 * R = build side (PK), S = probe side. Returns a malloc'd result_t the caller frees; when
 * result_type == 1, result->result is a chunked_table_t* owned by the caller
 * (destroy_table() + free(), as Join-Benchmarks/lib/TPCH-Queries/src/tpch.cpp:82 does). */
struct result_t *RHO(const struct table_t *relR, const struct table_t *relS, const struct joinconfig_t *config);

/* Join-Benchmarks/lib/Joins/src/ChunkedTable.cpp:128-136 */
void destroy_table(struct chunked_table_t *table);

/* Extra facts about the most recent join on this thread's context. The reference's result_t has
 * no checksum; checksum = sum over matches of (uint64)Rpayload + (uint64)Spayload
 * (CHT's convention, Joins/include/cht/CHTJoin.hpp:174), keysum = sum over matches of key. */
/* plan_flags: the join ran without a histogram pass (fixed-capacity partition regions) / it tried to,
 * a region overflowed (skew), and it was repeated with exact offsets - ms_* then describe the repeat only */
#define B200_PLAN_HISTOGRAM_FREE 1u
#define B200_PLAN_HISTOGRAM_FREE_OVERFLOWED 2u
#define B200_PLAN_HISTOGRAM_FREE_DECLINED 4u   /* a sample of the inputs showed skew: exact offsets from the start */
struct b200_join_stats_t {
    int64_t matches;
    uint64_t checksum;
    uint64_t keysum;
    uint32_t radix_bits;      /* total radix bits used */
    uint32_t num_passes;      /* partitioning passes (1 or 2) */
    uint32_t bits_pass1;
    uint32_t bits_pass2;
    uint32_t kernel_launches; /* kernels launched for this join */
    uint32_t plan_flags;      /* B200_PLAN_* */
    float ms_total;           /* device time, CUDA events: histogram .. build/probe */
    float ms_hist;
    float ms_pass1;
    float ms_pass2;
    float ms_join;
    float ms_h2d;             /* 0 for device-resident calls */
    float ms_d2h;
    float ms_materialize_host;
};
void b200_last_join_stats(struct b200_join_stats_t *out);

/* ------------------------------------------------------------------------------------------------
 * 2. join — ECALL-shaped split. Replaces ecall_preload_relations / ecall_join_preload
 *    (Join-Benchmarks/Enclave/Enclave.edl:46-53, secure_joins.cpp:34-58): copy once, run many.
 * ---------------------------------------------------------------------------------------------- */
int b200_preload_relations(const struct table_t *relR, const struct table_t *relS);
int b200_join_preload(const char *algorithm_name, const struct joinconfig_t *config, struct result_t *res);
void b200_free_preload(void);

/* ------------------------------------------------------------------------------------------------
 * 3. join — device-resident API (device pointers; stream = cudaStream_t or NULL for the
 *    library's own stream)
 * ---------------------------------------------------------------------------------------------- */
/* Whole local join on device-resident relations. If d_out != NULL up to out_capacity matches are
 * materialised as output_triple_t (order unspecified, like the reference's per-thread chunk
 * lists); stats->matches is always the full count. d_R / d_S are not modified. */
int b200_join_device(const struct row_t *d_R, uint64_t nR, const struct row_t *d_S, uint64_t nS,
                     struct output_triple_t *d_out, uint64_t out_capacity,
                     struct b200_join_stats_t *stats, void *stream);

/* Radix bits the GPU path picks for a build side of nR tuples (its analogue of
 * calc_num_radix_bits / calc_num_passes, radix_join.cpp:295-329, re-derived for shared memory). */
void b200_join_plan(uint64_t nR, uint32_t *total_bits, uint32_t *bits_pass1, uint32_t *bits_pass2);

/* Stage-level entry points (replace partition_hist :617-654, the prefix step :886-915 and
 * partition_copy :659-697 of radix_join.cpp). Used by the multi-GPU exchange path and by the
 * per-pass parity tests.
 *   hist    : d_hist[2^bits] (uint32) += count of digit ((key >> shift) & (2^bits-1)); caller zeroes.
 *   scatter : single-segment scatter of d_in[0..n) into d_out by digit. d_offsets[2^bits+1]
 *             (uint32, exclusive prefix of the histogram, on device) gives partition starts;
 *             d_cursors[2^bits] is scratch the call initialises from d_offsets. */
int b200_radix_hist_device(const struct row_t *d_in, uint64_t n, uint32_t shift, uint32_t bits,
                           uint32_t *d_hist, void *stream);
int b200_exclusive_scan_u32_device(const uint32_t *d_in, uint32_t n, uint32_t *d_out /* n+1 */, void *stream);
int b200_radix_scatter_device(const struct row_t *d_in, uint64_t n, uint32_t shift, uint32_t bits,
                              const uint32_t *d_offsets, uint32_t *d_cursors, struct row_t *d_out,
                              void *stream);

/* Sharded join across the GPUs of one box, one process per GPU. The first radix pass
 * doubles as the shuffle: a tuple is owned by the GPU whose rank equals the low log2_gpus bits of its
 * key, and pass 1 writes the partitions of each owner as one contiguous range, so its output is the
 * all-to-all send buffer (the host exchanges it with NCCL; this library does no communication).
 *
 * b200_shard_pass1_device — on every rank, for R and for S: histogram over total_bits key bits,
 *   pass-1 boundaries, scatter of d_in[0..n) into d_send by the ROUTED pass-1 digit
 *       p1' = rotate_right(key & (2^bits1-1), log2_gpus) within bits1 bits,
 *   so owner g's partitions are p1' in [g*2^bits1/G, (g+1)*2^bits1/G). Outputs: d_hist[2^total_bits]
 *   indexed by p1' | (p2 << bits1) (to be summed over ranks by the caller) and
 *   d_part1_off[2^bits1 + 1], the partition starts inside d_send.
 * b200_shard_join_device — on every rank after the exchange: the received relations consist of nseg
 *   segments (d_segoff_*[nseg+1]; segment s holds tuples of pass-1 partition group d_seg_group[s],
 *   several segments — one per source GPU — may share a group). Pass 2 scatters every segment by key
 *   bits [shift2, shift2+bits2) into its group's partitions, whose sizes d_hist_*[ngroups << bits2]
 *   (order group-major) are this rank's slice of the globally summed histograms; then build/probe
 *   with the hash on the key bits from hash_shift upward. stats receives this rank's partial
 *   matches / checksum / keysum (the caller sums them over ranks). */
int b200_shard_pass1_device(const struct row_t *d_in, uint64_t n, uint32_t total_bits, uint32_t bits1,
                            uint32_t log2_gpus, struct row_t *d_send, uint32_t *d_hist, uint32_t *d_part1_off,
                            void *stream);
/* Fused scatter + exchange (the B200 form of the shuffle): instead of writing a send buffer and handing
 * it to NCCL, the pass-1 scatter kernel stores every run directly into the receive buffer of the GPU
 * that owns the partition — peer memory mapped over NVLink 5 / NVSwitch — so the transfer overlaps the
 * partitioning tile by tile and the send buffer's HBM write + read disappear.
 *   b200_shard_hist_device    histogram of d_in[0..n) (as in b200_shard_pass1_device) + d_counts1[2^bits1],
 *                             this rank's size of every routed pass-1 partition. slot (0 = R, 1 = S) names
 *                             the workspace that carries per-CTA histogram rows to the scatter call.
 *   (host: all-gather the counts, derive d_dest_off[p] = where this rank's segment of partition p starts
 *    inside its owner's receive buffer)
 *   b200_shard_scatter_device scatters d_in into dest_bufs[owner(p)] + d_dest_off[p]; dest_bufs is a HOST
 *                             array of 2^log2_gpus device pointers (this GPU's own buffer and peers' buffers
 *                             opened with b200_ipc_open). The caller separates it from the readers of the
 *                             destination buffers with a cross-GPU barrier on the same stream.
 *   b200_ipc_*                export / open / close a buffer from b200_device_alloc across processes
 *                             (cudaIpc*; handle = 64 bytes). */
int b200_shard_hist_device(const struct row_t *d_in, uint64_t n, uint32_t total_bits, uint32_t bits1,
                           uint32_t log2_gpus, uint32_t *d_hist, uint32_t *d_counts1, int slot, void *stream);
int b200_shard_scatter_device(const struct row_t *d_in, uint64_t n, const uint32_t *d_dest_off, void *const *dest_bufs,
                              int slot, void *stream);
/* asynchronous device-to-device copy on `stream` (cudaMemcpyAsync, cudaMemcpyDefault): with a peer pointer
 * from b200_ipc_open as dst this is a copy-engine transfer over NVLink — the DMA form of the exchange */
int b200_copy_async(void *dst, const void *src, size_t bytes, void *stream);
int b200_ipc_export(void *d_ptr, unsigned char *handle_out /* 64 bytes */);
int b200_ipc_open(const unsigned char *handle /* 64 bytes */, void **d_ptr_out);
int b200_ipc_close(void *d_ptr);
int b200_shard_join_device(const struct row_t *d_R, uint64_t nR, const uint32_t *d_segoff_R, const struct row_t *d_S,
                           uint64_t nS, const uint32_t *d_segoff_S, const uint32_t *d_seg_group, uint32_t nseg,
                           uint32_t ngroups, uint32_t shift2, uint32_t bits2, const uint32_t *d_hist_R,
                           const uint32_t *d_hist_S, uint32_t hash_shift, struct b200_join_stats_t *stats,
                           void *stream);
/* Latency-trimmed forms for the fused exchange (the per-join fixed cost is what limits strong scaling at 8 GPUs):
 *   b200_exchange_plan_device     everything a rank derives from the two sizing collectives, in one launch.
 *       d_counts_all[world][2][2^bits1]  all-gathered d_counts1 of R and S;  d_hist_global[2][2^(bits1+bits2)] the
 *       all-reduced histograms. Outputs: d_seg_off[2][world*per+1] (per = 2^bits1/world; received segments ordered
 *       (source, local partition)), d_dest_off[2][2^bits1] (argument of b200_shard_scatter_device),
 *       d_hist_slice[2][per << bits2] (d_hist_R / d_hist_S of b200_shard_join_device) and d_host_vals[6] =
 *       {largest receive size of any rank R, S; this rank's receive sizes R, S; tuples this rank keeps R, S}.
 *   b200_shard_join_async_device  b200_shard_join_device without the host round trip: d_result3[3] receives
 *       {matches, checksum, keysum} on the device, in stream order; no synchronisation.
 *   b200_shard_join_times         phase times of the last (async) shard join; waits for its last event. */
int b200_exchange_plan_device(const uint32_t *d_counts_all, uint32_t world, uint32_t rank, uint32_t bits1,
                              uint32_t bits2, const uint32_t *d_hist_global, uint32_t *d_seg_off, uint32_t *d_dest_off,
                              uint32_t *d_hist_slice, uint64_t *d_host_vals, void *stream);
int b200_shard_join_async_device(const struct row_t *d_R, uint64_t nR, const uint32_t *d_segoff_R,
                                 const struct row_t *d_S, uint64_t nS, const uint32_t *d_segoff_S,
                                 const uint32_t *d_seg_group, uint32_t nseg, uint32_t ngroups, uint32_t shift2,
                                 uint32_t bits2, const uint32_t *d_hist_R, const uint32_t *d_hist_S, uint32_t hash_shift,
                                 uint64_t *d_result3, void *stream);
int b200_shard_join_times(struct b200_join_stats_t *stats);

/* Multi-GPU host of the sharded join (csrc/mg.cu): the C form of what join_init_run does for threads
 * (radix_join.cpp:1369-1638), for G = 1, 2, 4 or 8 processes of one box, one per GPU. Every process calls
 *   b200_init(local device); b200_mg_init(rank, world, id, |R| total, |S| total)     once
 *   b200_mg_join(its row-range shard of R, of S, &result)                            per join (collective)
 *   b200_mg_finalize()                                                              once (collective)
 * `id` is the 128-byte NCCL unique id rank 0 obtains from b200_mg_unique_id and hands to the others by any means
 * (the C driver host/native_mg.cpp uses shared memory between forked processes, bench.py torch.distributed).
 * NCCL carries three small collectives per join (partition counts, histograms, the 24-byte result) and one
 * barrier; the tuples move inside the pass-1 scatter kernel, stored straight into the owners' buffers over NVLink
 * peer memory (CUDA IPC). A shard may hold at most ceil(total / world) tuples. result: GLOBAL matches / checksum /
 * keysum on every rank, this rank's phase times (CUDA events). */
struct b200_mg_result_t {
    uint64_t matches, checksum, keysum;
    uint64_t tuples_sent;          /* tuples this rank scattered (its shard) */
    uint64_t tuples_kept;          /* of those, tuples whose owner is this rank (never cross NVLink) */
    uint32_t radix_bits, bits_pass1, bits_pass2, world, kernel_launches, reserved;
    float ms_total;                /* first histogram .. global result on the device */
    float ms_hist, ms_scatter, ms_barrier, ms_local, ms_reduce;   /* consecutive phases of ms_total */
    float ms_pass2, ms_join;       /* inside ms_local */
};
int b200_mg_unique_id(unsigned char *id_out /* 128 bytes */);
int b200_mg_init(int rank, int world, const unsigned char *id /* 128 bytes */, uint64_t nR_total, uint64_t nS_total);
/* same with explicit limits: capR / capS = the largest shard any rank will pass to b200_mg_join (relations whose shards
 * are uneven, e.g. the output of a row-range filter), dead_bits = low key bits known to carry no information (TPC-H
 * order keys: 2), which the radix plan then skips as the single-GPU planner does */
int b200_mg_init_caps(int rank, int world, const unsigned char *id /* 128 bytes */, uint64_t nR_total, uint64_t capR,
                      uint64_t capS, uint32_t dead_bits);
int b200_mg_join(const struct row_t *d_R, uint64_t nR_local, const struct row_t *d_S, uint64_t nS_local,
                 struct b200_mg_result_t *result);
/* Materialising form (the MATERIALIZE switch of joinconfig_t, radix_join.cpp:428-447,:1556, for sharded relations): the
 * triples {key, Rpayload, Spayload} of the co-partitions THIS rank owns are left in a library-owned device buffer
 * (*d_triples, *local_rows of them; valid until the next b200_mg_* call) - the result stays sharded by key, ready to be
 * the input of the next sharded join. result->matches is the global count. The buffer is sized for a unique build
 * key and even shards; if this rank produces more, it is grown and the probe (only) runs again. */
int b200_mg_join_materialize(const struct row_t *d_R, uint64_t nR_local, const struct row_t *d_S, uint64_t nS_local,
                             const struct output_triple_t **d_triples, uint64_t *local_rows,
                             struct b200_mg_result_t *result);
/* Sum of n (1..4) host values over all ranks, in place (collective): turns per-rank counts of a sharded pipeline into
 * its answer without a second communication layer. */
int b200_mg_allreduce_u64(uint64_t *values, int n);
int b200_mg_finalize(void);

/* ------------------------------------------------------------------------------------------------
 * 4. relation generators
 * ---------------------------------------------------------------------------------------------- */
/* Host generators, bit-identical to the reference for the same seed (they run the same libc
 * srand()/rand() sequence): Join-Benchmarks/lib/AppUtilities/include/generator.h:26-107,
 * src/generator.cpp:75,:352,:474,:515,:638,:663. tuples are malloc'd; payload is set to 0
 * (the reference leaves it uninitialised). Return 0 on success. */
void seed_generator(unsigned int seed);
int create_relation_pk(struct table_t *reln, uint64_t ntuples, int sorted);
int create_relation_fk(struct table_t *reln, uint64_t ntuples, const int64_t maxid, int sorted);
int create_relation_fk_sel(struct table_t *reln, uint64_t ntuples, const int64_t maxid, int sorted);
/* The reference seeds Zipf from std::random_device (genzipf.cpp:44-45,:104-105); this one draws its
 * std::mt19937_64 seed from the seed_generator() value so runs are repeatable. Same LUT + binary
 * search algorithm (genzipf.cpp:58-137). */
int create_relation_zipf(struct table_t *reln, uint64_t ntuples, const int64_t maxid, const double zipfparam,
                         int sorted);
void delete_relation(struct table_t *reln);

/* Device generators ("gpu" mode): write tuples [row_begin, row_begin+n) of a
 * relation of n_total tuples straight into HBM at d_rel[0..n). Same key *distribution* as the
 * reference (PK: a permutation of 1..n_total; FK: floor(n_total/maxid) independent permutations of
 * 1..maxid laid end to end, + remainder 1..rem), different permutation (counter-based bijection
 * instead of the sequential glibc Knuth shuffle). payload = global row index. */
int b200_gen_pk_device(struct row_t *d_rel, uint64_t n_total, uint64_t row_begin, uint64_t n,
                       uint64_t seed, void *stream);
int b200_gen_fk_device(struct row_t *d_rel, uint64_t n_total, uint64_t maxid, uint64_t row_begin, uint64_t n,
                       uint64_t seed, void *stream);
/* Zipf-skewed foreign keys (create_relation_zipf, generator.cpp:638-660 -> gen_zipf, genzipf.cpp:87-144):
 * keys drawn from a random permutation ("alphabet") of 1..maxid with probability proportional to
 * rank^-zipf_param, by binary search in the cumulative table the reference builds. The table (8*maxid
 * bytes, temporary) is computed on the device; rows [row_begin, row_begin+n) of the stream are written. */
int b200_gen_zipf_device(struct row_t *d_rel, uint64_t maxid, double zipf_param, uint64_t row_begin, uint64_t n,
                         uint64_t seed, void *stream);
/* payload[i] = row_begin + i for an existing device relation (TPC-H loader convention,
 * Join-Benchmarks/App/TpcH/TpcHCommons.cpp:332,:413) */
int b200_set_rowid_payload_device(struct row_t *d_rel, uint64_t row_begin, uint64_t n, void *stream);

/* ------------------------------------------------------------------------------------------------
 * 5. scans over a packed uint8 column: predicate lo <= v <= hi (unsigned, inclusive); only
 *    num_records/64 whole 64-value blocks are processed, like every SIMD512 loop
 *    (Scan-Micro-Benchmarks/shared_libraries/SimdScan/src/SIMD512.cpp:216,:234,:264).
 * ---------------------------------------------------------------------------------------------- */
/* Replaces ecall_bitvector_scan_user (SimdScanMulti/Enclave/Enclave.edl:71-79, Enclave.cpp:270-299)
 * = SIMD512::bitvector_scan (SIMD512.cpp:210-222). Host buffers: data[num_records] and
 * output_buffer[num_records/64]. Does warmup_runs untimed + num_runs timed passes over
 * device-resident data (unique_data != 0 forces 1 run / 0 warm-ups) and ADDS the timed device
 * nanoseconds to *time_cntr (the reference adds TSC cycles to *cpu_cntr). The H2D/D2H copies are
 * outside that counter (the reference's data is already in memory); see b200_scan_last_copy_ns. */
void b200_bitvector_scan_user(uint8_t predicate_low, uint8_t predicate_high, const uint8_t *data,
                              size_t num_records, uint64_t *output_buffer, uint64_t *time_cntr,
                              size_t num_runs, size_t warmup_runs, int unique_data);

/* Replaces ecall_index_scan_user (Enclave.edl:33-41, Enclave.cpp:100-133) =
 * SIMD512::implicit_index_scan_self_alloc (SIMD512.cpp:251-287) with a plain output array instead
 * of a CacheAlignedVector*: writes the ascending uint64 positions (relative to `data`) of matching
 * values to output_buffer[0..min(count, output_capacity)) and the match count to *output_count. */
void b200_index_scan_user(uint8_t predicate_low, uint8_t predicate_high, const uint8_t *data,
                          size_t num_records, uint64_t *output_buffer, size_t output_capacity,
                          size_t *output_count, uint64_t *time_cntr, size_t num_runs, size_t warmup_runs,
                          int unique_data);
/* H2D + D2H nanoseconds (host clock) of the most recent *_user scan call */
uint64_t b200_scan_last_copy_ns(void);

/* Device-resident forms. d_data must be 16-byte aligned. */
int b200_bitvector_scan_device(uint8_t lo, uint8_t hi, const uint8_t *d_data, size_t num_records,
                               uint64_t *d_out, void *stream);
/* SIMD512::count (SIMD512.cpp:7-32); result written to *d_count (device uint64) */
int b200_scan_count_device(uint8_t lo, uint8_t hi, const uint8_t *d_data, size_t num_records,
                           uint64_t *d_count, void *stream);
/* row ids = id_base + position; *d_count (device uint64) receives the match count; ids beyond
 * out_capacity are dropped (count is still exact). */
int b200_index_scan_device(uint8_t lo, uint8_t hi, const uint8_t *d_data, size_t num_records,
                           uint64_t id_base, uint64_t *d_out_ids, uint64_t out_capacity,
                           uint64_t *d_count, void *stream);
/* The remaining SIMD512 variants on the 8-bit column. They share the row-id machinery
 * (bitvector + per-tile counts -> offsets -> expansion); only what is written per match differs.
 *   b200_scan_sum_device               SIMD512::sum  (SIMD512.cpp:34-86): sum of the values in range -> *d_sum
 *   b200_value_scan_device             SIMD512::scan (:89-150): the matching values as uint32, in input order
 *   b200_dict_scan_8bit_64bit_device   SIMD512::dict_scan_8bit_64bit (:289-336): dict[code] (int64) of every code whose
 *       dictionary value lies in [predicate_low, predicate_high]. dict = 256 sorted int64 in HOST memory (it is
 *       searched on the host exactly like the reference's two std::find_if calls, :297-305, including the uint8
 *       wrap-around for predicates outside the dictionary's range, then copied — 2 KiB — with the call).
 * d_count (device uint64) receives the exact match count; values beyond out_capacity are dropped. */
int b200_scan_sum_device(uint8_t lo, uint8_t hi, const uint8_t *d_data, size_t num_records, uint64_t *d_sum, void *stream);
int b200_value_scan_device(uint8_t lo, uint8_t hi, const uint8_t *d_data, size_t num_records, uint32_t *d_out,
                           uint64_t out_capacity, uint64_t *d_count, void *stream);
int b200_dict_scan_8bit_64bit_device(int64_t predicate_low, int64_t predicate_high, const int64_t *dict,
                                     const uint8_t *d_data, size_t num_records, int64_t *d_out, uint64_t out_capacity,
                                     uint64_t *d_count, void *stream);
/* Host-buffer forms in the reference functions' argument order (SIMD512.hpp:39-84) with a plain output array and its
 * capacity instead of a CacheAlignedVector&; they return the sum resp. the exact match count and abort like the
 * reference's allocation failures do (util.cpp:12-19) if the device is unusable. */
uint64_t b200_sum(uint8_t predicate_low, uint8_t predicate_high, const uint8_t *data, size_t num_records);
uint64_t b200_scan(uint8_t predicate_low, uint8_t predicate_high, const uint8_t *data, size_t num_records,
                   uint32_t *output_buffer, size_t output_capacity);
uint64_t b200_dict_scan_8bit_64bit(int64_t predicate_low, int64_t predicate_high, const int64_t *dict, const uint8_t *data,
                                   size_t num_records, int64_t *output_buffer, size_t output_capacity);
/* The remaining SIMD512 scans, reference argument order, plain output array + capacity in place of the
 * CacheAlignedVector; every call returns the exact number of matches (also when it exceeds the capacity).
 *   explicit_index_scan  SIMD512.cpp:152-208: like the row-id scan, but a match at position p emits the caller's index entry
 *                        index[((p / 64) + (p / 8) % 8) * 8 + p % 8] - block i's byte group j reads index register i + j,
 *                        exactly as the reference indexes it; `index` holds (n / 64 + 7) * 8 entries.
 *   scalar_index_scan    ScalarScan.hpp:8-20: the scalar twin of the row-id scan - ALL n values, not only n / 64 blocks.
 *   dict_scan_16bit / 32bit_64bit  SIMD512.cpp:531-622: dict[code] of every code whose dictionary value lies in
 *                        [predicate_low, predicate_high]; 65536-entry dictionary resp. dict_size entries; whole 512-bit
 *                        registers only (32 resp. 16 codes); the code range goes through uint16_t as in the reference. */
uint64_t b200_explicit_index_scan(uint8_t lo, uint8_t hi, const uint64_t *index, const uint8_t *data, size_t n,
                                  uint64_t *output_buffer, size_t output_capacity);
int b200_explicit_index_scan_device(uint8_t lo, uint8_t hi, const uint64_t *d_index, const uint8_t *d_data, size_t n,
                                    uint64_t *d_out, uint64_t out_capacity, uint64_t *d_count, void *stream);
uint64_t b200_scalar_index_scan(uint8_t lo, uint8_t hi, const uint8_t *data, size_t n, uint64_t *output_buffer,
                                size_t output_capacity);
uint64_t b200_dict_scan_16bit_64bit(int64_t predicate_low, int64_t predicate_high, const int64_t *dict, const uint16_t *data,
                                    size_t n, int64_t *output_buffer, size_t output_capacity);
uint64_t b200_dict_scan_32bit_64bit(int64_t predicate_low, int64_t predicate_high, const int64_t *dict, size_t dict_size,
                                    const uint32_t *data, size_t n, int64_t *output_buffer, size_t output_capacity);
/* device form of both: code_bits = 16 or 32, the code range already derived, dictionary and column in device memory */
int b200_dict_scan_wide_device(int code_bits, uint32_t code_lo, uint32_t code_hi, const int64_t *d_dict, const void *d_data,
                               size_t n, int64_t *d_out, uint64_t out_capacity, uint64_t *d_count, void *stream);

/* Allocator.hpp:95-109 tiled 0..255 column, generated in HBM; value at global position p is p mod 256 */
int b200_fill_tiled_column_device(uint8_t *d_data, size_t n, uint64_t pos_begin, void *stream);
/* seeded skewed column for selectivities the tiled column cannot express:
 * v = 0 with probability p_zero_ppm / 1e6, else uniform in 1..255 */
int b200_fill_skewed_column_device(uint8_t *d_data, size_t n, uint64_t pos_begin, uint32_t p_zero_ppm,
                                   uint64_t seed, void *stream);

/* number of CUDA kernels this library launched since load (bench.py reports the delta) */
uint64_t b200_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* AQP_B200_AQP_H */
