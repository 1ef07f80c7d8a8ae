#!/usr/bin/env python
"""bench.py — the hot path's headline benchmark (BASELINE.json): RHO radix hash join throughput in
(|R|+|S|) Mtuples/s with the scan's GB/s beside it, against the HBM roofline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one complete join (histogram -> prefix sums -> pass-1 scatter -> pass-2 scatter ->
build/probe) of device-resident relations |R|=2^27, |S|=2^29 (BASELINE config 3; 5 GiB of input, far
larger than the 126 MB L2, so no cache flush is needed between steps). Prints ONE JSON line.

  value     join throughput, inputs resident in HBM, device time (CUDA events), max over ranks
  e2e       same join through the drop-in C ABI call run_join(result_t*, R, S, "RHO", cfg) with
            pinned HOST relations: H2D of 8(|R|+|S|) bytes and the result read-back are inside the
            timed region
  roofline  the dominant kernel (radix_scatter_bins_kernel, 4 launches per join): algorithmic bytes per
            launch (16 B/tuple) / its mean launch time from the library's CUDA events
  join_roofline  the whole join at SURVEY.md §8d's graded 56 B/tuple
  scan      bitvector and row-id scans over 2^30 uint8 (BASELINE config 2), GB/s of input and
            fraction of the HBM peak at 1.125 B/value resp. (1 + 8 sel) B/value
  cpu_baseline   the reference's own RHO (oracle/_ref, compiled from /root/reference) on all of this box's
            host cores at the SAME size 2^27 x 2^29 (a few runs) — a reported baseline, not the target
  skew / tpch / exchange   BASELINE configs 4 and 5, NVLink traffic of the shuffle at N > 1

--impl reference runs only that CPU reference arm and prints the same line shape.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "sgxv2-analytical-query-processing-benchmarks_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

LOG_R, LOG_S = 27, 29          # BASELINE config 3
SCAN_LOG_N = 30                # BASELINE config 2
JOIN_BYTES_PER_TUPLE = 56      # 8 * (3 P + 1) with P = 2 passes
SCATTER_BYTES_PER_TUPLE = 16   # read 8 + write 8
# multi-GPU shuffle: "mg" (the C host in csrc/mg.cu: scatter fused with peer stores over NVLink, region layout),
# "p2p" / "dma" / "nccl" (round-1 Python-orchestrated forms: fused with sizing collectives in front, copy engines,
# all_to_all_single). Measured ms per join at 2 GPUs in one run (profiles/r02_bench_2gpu_a.json): mg 4.38, p2p 4.49,
# dma 6.08, nccl 8.82; at 8 GPUs mg 1.86 (p2p in round 1: 2.00) -> mg everywhere
DEFAULT_EXCHANGE = {}
METRIC = "rho_join_throughput"
UNIT = "Mtuples/s"


def workload_config(world: int) -> dict:
    """`config` of the JSON line - the same dict in the B200 arm and in the reference arm (same workload, same sizes)."""
    return {"workload": f"RHO join |R|=2^{LOG_R} |S|=2^{LOG_S} 8-byte tuples (u32 key, u32 payload), uniform FK keys, "
                        "count+checksum (BASELINE config 3)",
            "cache": "inputs (5 GiB) exceed every cache (126 MB L2); no flush between steps",
            "parallelism": f"{world} GPU(s), one process per GPU"}


def host_info() -> dict:
    """nproc, CPU model and the measured TSC rate (SURVEY.md 8d asks for all three beside a CPU number)."""
    info = {"nproc": os.cpu_count()}
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                info["cpu_model"] = ln.split(":", 1)[1].strip()
                break
    except OSError:
        pass
    return info


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"   # B200_PROFILING.md fallback


def ncu_traffic(kernel_prefix: str):
    """DRAM bytes per launch of the dominant kernel (mean over its captured launches and template instances) from
    the committed `ncu --set full` capture (profiles/traffic.json, written by tools/ncu_traffic.py); None if there
    is no capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        n, b = 0, 0.0
        for name, e in t.items():
            if name.startswith(kernel_prefix):
                n += e["launches"]
                b += e["dram_bytes_per_launch"] * e["launches"]
        return b / n if n else None
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# CPU reference arm (oracle/_ref = the unmodified reference compiled from /root/reference)
# --------------------------------------------------------------------------------------------------
def _fast_inputs(log_r: int, log_s: int):
    """PK / FK relations with the reference generator's distribution (R = a permutation of 1..|R|, S = |S|/|R|
    independent permutations laid end to end, payload = row id; generator.cpp:143-153,:474-512) from numpy
    permutations: the reference's own sequential glibc shuffle needs ~1 minute at 2^29 and none of it is timed."""
    import numpy as np
    import oracle as O
    nR, nS = 1 << log_r, 1 << log_s
    rng = np.random.default_rng(11111)
    R = np.empty(nR, dtype=O.ROW)
    R["key"] = rng.permutation(nR).astype(np.uint32) + 1
    R["payload"] = np.arange(nR, dtype=np.uint32)
    S = np.empty(nS, dtype=O.ROW)
    rng = np.random.default_rng(22222)
    for c in range(nS // nR):
        S["key"][c * nR:(c + 1) * nR] = rng.permutation(nR).astype(np.uint32) + 1
    S["payload"] = np.arange(nS, dtype=np.uint32)
    return R, S


def cpu_reference_join(steps: int, warmup: int, log_r: int = LOG_R, log_s: int = LOG_S):
    """The reference's own RHO (oracle/_ref: the unmodified sources compiled from /root/reference) on all host cores
    at the headline size 2^27 x 2^29 (falls back to 2^24 x 2^26 only if the host cannot hold it, and says so).
    Throughput = tuples / (the reference's own 'Total Join Time (cycles)' / measured TSC rate); the
    wall-clock of the RHO() call (which also covers its buffer allocation) is reported beside it. Falls back to the
    oracle port if oracle/_ref cannot run here (no AVX-512)."""
    import oracle as O
    cores = os.cpu_count() or 1
    note = ""
    t0 = time.time()
    try:
        R, S = _fast_inputs(log_r, log_s)
    except MemoryError:
        log_r, log_s = 24, 26
        note = " (host memory too small for 2^27 x 2^29: 1/8 sample)"
        R, S = _fast_inputs(log_r, log_s)
    nR, nS = 1 << log_r, 1 << log_s
    gen_s = time.time() - t0
    kind = "reference" if O.have_ref() else "port"
    times, walls = [], []
    matches = None
    best_flags = None
    tsc = None
    if kind == "reference":
        # paper-best flags UNROLL+FORCE_2_PHASES and the automatic pass count; keep the better
        probe = {}
        for force2 in (True, False):
            r = O.ref_rho(R, S, nthreads=cores, force_2_passes=force2)
            probe[force2] = r["seconds"]
        best_flags = min(probe, key=probe.get)
        for i in range(warmup + steps):
            r = O.ref_rho(R, S, nthreads=cores, force_2_passes=best_flags)
            matches = r["matches"]
            tsc = r["tsc_hz"]
            if i >= warmup:
                times.append(r["seconds"])
                walls.append(r["wall_seconds"])
    else:
        cores = 1
        for i in range(max(1, min(steps, 3))):
            t = time.time()
            r = O.rho(R, S, nthreads=1)
            times.append(time.time() - t)
            walls.append(times[-1])
            matches = r["matches"]
    assert matches == nS
    mean = sum(times) / len(times)
    return {"value": (nR + nS) / mean / 1e6, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"|R|=2^{log_r} |S|=2^{log_s} uniform FK{note}, {len(times)} runs after {warmup} warm-ups, "
                      f"flags UNROLL{'+FORCE_2_PHASES' if best_flags else ''}, numpy-permutation inputs "
                      f"({gen_s:.1f}s, untimed)",
            "ms_per_step": mean * 1e3, "wall_ms_per_step": sum(walls) / len(walls) * 1e3, "tsc_hz": tsc,
            "same_size_as_headline": (log_r, log_s) == (LOG_R, LOG_S), "host": host_info()}


def cpu_reference_scan(log_n: int = 28):
    import numpy as np
    import oracle as O
    if not O.have_ref():
        return None
    n = 1 << log_n
    cores = os.cpu_count() or 1
    col = O.aligned_u8(n)
    col[:] = O.tiled_column(n)
    out = np.zeros(n // 64, dtype=np.uint64)
    L = O.ref_scan()
    res = {"cores": cores, "kind": "reference", "sample": f"2^{log_n} tiled uint8, 3 runs after 1 warm-up"}
    s = L.ref_scan_mt(0, 0, 26, col.ctypes.data, n, cores, 1, 3, out.ctypes.data)
    res["bitvector_gbs"] = n * 3 / s / 1e9
    for name, hi in (("rowid_sel10_gbs", 26), ("rowid_sel100_gbs", 255)):
        s = L.ref_scan_mt(1, 0, hi, col.ctypes.data, n, cores, 1, 3, None)
        res[name] = n * 3 / s / 1e9
    return res


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_reference_join(args.steps, args.warmup)
    scan = cpu_reference_scan()
    keep = ("value", "unit", "cores", "kind", "sample", "wall_ms_per_step", "tsc_hz", "same_size_as_headline", "host")
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": workload_config(args.gpus),
            "cpu_baseline": {k: cb[k] for k in keep},
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "scan": scan}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------
def bench_scan(A, torch, dev, peak, steps, warmup, n_gpus=1, rank=0):
    """Row-range shard of the 2^30 column on this rank (trivially parallel, no exchange)."""
    n_total = 1 << SCAN_LOG_N
    n = n_total // n_gpus
    begin = rank * n
    data = torch.empty(n, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    assert A.lib().b200_fill_tiled_column_device(data.data_ptr(), n, begin, A._st(st)) == 0
    bv = torch.empty(n // 64, dtype=torch.int64, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    out = {}
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timeit(fn):
        for _ in range(max(warmup, 3)):
            fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    ms = timeit(lambda: A.bitvector_scan_device(0, 26, data.data_ptr(), n, bv.data_ptr(), st))
    out["bitvector"] = {"predicate": [0, 26], "ms": ms, "input_gbs": n / ms / 1e6,
                        "roofline": {"bound": "hbm", "achieved": 1.125 * n / ms / 1e6, "peak": peak, "unit": "GB/s",
                                     "frac": 1.125 * n / ms / 1e6 / peak}}
    # 0.39 % / 10.5 % / 50.4 % / 100 % true selectivity on the tiled column (types.hpp:125 mapping)
    for label, hi in (("sel0.4", 0), ("sel10", 26), ("sel50", 128), ("sel100", 255)):
        k = n // 256 * (hi + 1)
        ids = torch.empty(k, dtype=torch.int64, device=dev)
        ms = timeit(lambda: A.index_scan_device(0, hi, data.data_ptr(), n, ids.data_ptr(), k, cnt.data_ptr(),
                                                id_base=begin, stream=st))
        assert int(cnt.item()) == k
        alg = n + 8 * k
        out["rowid_" + label] = {"predicate": [0, hi], "selectivity": k / n, "ms": ms, "input_gbs": n / ms / 1e6,
                                 "roofline": {"bound": "hbm", "achieved": alg / ms / 1e6, "peak": peak, "unit": "GB/s",
                                              "frac": alg / ms / 1e6 / peak}}
        del ids
    # BASELINE config 2's 0.1 %: not expressible on the tiled column -> seeded skewed column, v = 0 with
    # probability 1e-3 else uniform 1..255, predicate [0, 0]; the true selectivity is counted, not assumed
    assert A.lib().b200_fill_skewed_column_device(data.data_ptr(), n, begin, 1000, 42, A._st(st)) == 0
    A.scan_count_device(0, 0, data.data_ptr(), n, cnt.data_ptr(), st)
    torch.cuda.synchronize()
    k = int(cnt.item())
    ids = torch.empty(max(k, 1), dtype=torch.int64, device=dev)
    ms = timeit(lambda: A.index_scan_device(0, 0, data.data_ptr(), n, ids.data_ptr(), k, cnt.data_ptr(),
                                            id_base=begin, stream=st))
    assert int(cnt.item()) == k
    alg = n + 8 * k
    out["rowid_sel0.1"] = {"predicate": [0, 0], "column": "skewed: 0 w.p. 1e-3 else uniform 1..255, seed 42",
                           "selectivity": k / n, "ms": ms, "input_gbs": n / ms / 1e6,
                           "roofline": {"bound": "hbm", "achieved": alg / ms / 1e6, "peak": peak, "unit": "GB/s",
                                        "frac": alg / ms / 1e6 / peak}}
    out["n_values_per_gpu"] = n
    # DRAM bytes per launch from the committed `ncu --set full` capture of the same kernels at 2^30 values
    # (profiles/r02_traffic_scan.json from tools/scan_probe.py: skewed, then the tiled column at hi = 0, 26, 128, 255)
    if n == 1 << SCAN_LOG_N:
        try:
            with open(os.path.join(ROOT, "profiles", "r02_traffic_scan.json")) as f:
                t = json.load(f)
            fused = next(v for k, v in t.items() if k.startswith("aqp::rowid_scan_fused_kernel") or "rowid_scan_fused_kernel" in k)
            for key, i in (("rowid_sel0.1", 0), ("rowid_sel0.4", 1), ("rowid_sel10", 2), ("rowid_sel50", 3), ("rowid_sel100", 4)):
                out[key]["roofline"]["traffic"] = fused["per_launch"][i]
            bvk = next(v for k, v in t.items() if "bitvector_scan_tma_kernel" in k)
            out["bitvector"]["roofline"]["traffic"] = bvk["per_launch"][0]
        except Exception:
            pass
    del ids, data, bv
    return out


def bench_scan_e2e(A, np_mod, steps):
    """Scan end to end through the ECALL-shaped host-buffer calls (b200_bitvector_scan_user / b200_index_scan_user,
    Enclave.edl:33-41,:71-79): H2D of the column, the kernel, D2H of the bitvector / id list inside the timed region."""
    import torch
    n = 1 << SCAN_LOG_N
    col = torch.empty(n, dtype=torch.uint8).pin_memory()
    col.copy_((torch.arange(n, dtype=torch.int32) & 255).to(torch.uint8))
    npcol = col.numpy()
    res = {}
    for name, fn, d2h in (("bitvector", lambda: A.bitvector_scan_user(0, 26, npcol), n // 8),
                          ("rowid_sel10", lambda: A.index_scan_user(0, 26, npcol, capacity=n // 256 * 27 + 64), n // 256 * 27 * 8)):
        fn()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        dt = (time.perf_counter() - t0) / steps
        res[name] = {"ms": dt * 1e3, "input_gbs": n / dt / 1e9, "h2d_bytes_per_step": n, "d2h_bytes_per_step": d2h}
    res["api"] = "b200_bitvector_scan_user / b200_index_scan_user on a pinned host column (output arrays pageable numpy)"
    return res


def run_b200_arm(args):
    import torch
    import b200aqp as A

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libb200aqp has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    A.init(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        import b200aqp.dist as D
    peak, peak_src = measured_peaks()
    nR, nS = 1 << LOG_R, 1 << LOG_S
    st = torch.cuda.current_stream().cuda_stream
    launches0 = A.kernel_launch_count()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- inputs generated straight into HBM (row-range shard per rank) ------------------------------
    nR_loc, nS_loc = nR // world, nS // world
    R = torch.empty(nR_loc * 2, dtype=torch.int32, device=dev)
    S = torch.empty(nS_loc * 2, dtype=torch.int32, device=dev)
    A.gen_pk_device(R.data_ptr(), nR, 11111, rank * nR_loc, nR_loc, st)
    A.gen_fk_device(S.data_ptr(), nS, nR, 22222, rank * nS_loc, nS_loc, st)
    torch.cuda.synchronize()

    if world == 1:
        def step():
            return A.join_device(R.data_ptr(), nR, S.data_ptr(), nS, stream=st)
    else:
        # headline: scatter kernel fused with the exchange over NVLink peer memory; the NCCL all-to-all
        # variant is timed beside it as the baseline (B200_AQP_EXCHANGE=nccl makes it the headline)
        # "mg" = the C host inside the library (csrc/mg.cu, what host/native_mg.cpp drives): region-layout exchange, the
        # sizing collectives travel beside the scatter. The round-1 Python-orchestrated variants stay for A/B.
        variants = {"mg": D.MgShardedJoin, "p2p": D.FusedShardedJoin, "dma": D.DmaShardedJoin, "nccl": D.ShardedJoin}
        default = DEFAULT_EXCHANGE.get(world, "mg")
        plan = variants[os.environ.get("B200_AQP_EXCHANGE", default)](nR, nS, dev)

        def step():
            return plan.run(R, S)

    for _ in range(max(args.warmup, 3)):
        s = step()
    assert s["matches"] == nS, s
    rep = nS // nR
    assert s["keysum"] == rep * nR * (nR + 1) // 2 and s["checksum"] == rep * (nR * (nR - 1) // 2) + nS * (nS - 1) // 2

    sampler = ClockSampler(local_rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    phase = {"ms_hist": 0.0, "ms_pass1": 0.0, "ms_pass2": 0.0, "ms_join": 0.0, "ms_total": 0.0, "ms_exchange": 0.0,
             "ms_sizing": 0.0, "ms_scatter_kernels": 0.0, "ms_barrier": 0.0}
    l_before = A.kernel_launch_count()
    barrier()
    sampler.start()
    e0.record()
    for _ in range(args.steps):
        s = step()
        for k in phase:
            phase[k] += s.get(k, 0.0)
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches_timed = A.kernel_launch_count() - l_before
    ms_step = e0.elapsed_time(e1) / args.steps
    if world > 1:
        t = torch.tensor([ms_step], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step = float(t.item())
    for k in phase:
        phase[k] /= args.steps
    value = (nR + nS) / ms_step / 1e3   # Mtuples/s

    # dominant kernel: radix_scatter_bins_kernel, 4 launches per join (R and S, pass 1 and pass 2)
    scatter_ms = (phase["ms_pass1"] + phase["ms_pass2"]) / 4
    scatter_bytes = SCATTER_BYTES_PER_TUPLE * (nR_loc + nS_loc) * 2 / 4
    roof = {"bound": "hbm", "kernel": "radix_scatter_bins_kernel", "achieved": scatter_bytes / scatter_ms / 1e6 if scatter_ms else None,
            "peak": peak, "peak_source": peak_src, "unit": "GB/s",
            "traffic": ncu_traffic("radix_scatter_bins_kernel") if world == 1 else None,
            "bytes_per_launch": scatter_bytes, "ms_per_launch": scatter_ms}
    roof["frac"] = roof["achieved"] / peak if roof["achieved"] else None
    join_bytes = JOIN_BYTES_PER_TUPLE * (nR + nS) / world
    join_roof = {"bound": "hbm", "bytes_per_tuple": JOIN_BYTES_PER_TUPLE, "achieved": join_bytes / ms_step / 1e6,
                 "peak": peak, "unit": "GB/s", "frac": join_bytes / ms_step / 1e6 / peak,
                 "actual_bytes_per_tuple": 48, "note": "the pass-2 histogram read is avoided: one full-width histogram serves both passes"}
    if s.get("plan_flags", 0) & 1:   # B200_PLAN_HISTOGRAM_FREE: partitions live in fixed-capacity regions, no histogram pass at all
        join_roof["actual_bytes_per_tuple"] = 40
        join_roof["note"] = ("histogram-free plan (fixed-capacity partition regions, exact-offset repeat on overflow): no histogram "
                             "read at all, 2 x 16 B/tuple scatter + 8 B/tuple build/probe; frac is still quoted at the graded 56 B/tuple")

    # ---- BASELINE config 4: Zipf-skewed S (z = 0.5, 1.0), same sizes, 1 GPU ----------------------------
    skew = {}
    for z in (0.5, 1.0):
        A.gen_zipf_device(S.data_ptr(), nS_loc, nR, z, 22222, rank * nS_loc, st)   # this rank's rows of the Zipf stream
        torch.cuda.synchronize()
        for _ in range(2):
            sz = step()
        assert sz["matches"] == nS, sz          # every Zipf key is in 1..|R| and R is a primary key
        barrier()
        e0.record()
        nz = max(3, args.steps // 2)
        for _ in range(nz):
            sz = step()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / nz
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        skew[f"z{z}"] = {"ms": ms, "value": (nR + nS) / ms / 1e3, "unit": UNIT, "checksum": sz["checksum"],
                         "join_roofline_frac": JOIN_BYTES_PER_TUPLE * (nR + nS) / world / ms / 1e6 / peak}
    A.gen_fk_device(S.data_ptr(), nS, nR, 22222, rank * nS_loc, nS_loc, st)   # restore the uniform S
    torch.cuda.synchronize()

    # ---- multi-GPU: NVLink traffic of the shuffle, and the NCCL all-to-all variant as the baseline -------
    exchange = None
    if world > 1:
        sent = 8 * ((nR_loc + nS_loc) - s.get("tuples_kept", 0))
        kind = s.get("exchange", "nccl all_to_all_single")
        ms_x = phase["ms_exchange"] if kind.startswith("nccl") else phase["ms_pass1"]
        exchange = {"kind": kind, "bytes_sent_per_gpu": sent, "ms": ms_x, "busbw_gbs": sent / ms_x / 1e6 if ms_x else None,
                    "note": "p2p-fused: ms = sizing collectives + fused scatter/exchange kernel + barrier; p2p-dma: ms = local "
                            "pass-1 scatter (sizing hidden underneath) + copy-engine transfers + barrier; nccl: sizing + "
                            "all_to_all_single. Reference peaks: 770 GB/s measured peer copy, 900 GB/s nominal per direction",
                    "variants": {}}
        # the other exchange implementations timed beside the headline one (same inputs, same step count)
        # (opt-in: B200_AQP_BENCH_VARIANTS=p2p,dma,nccl - the round-1 Python-orchestrated forms; their numbers are in
        # profiles/r01_bench_{2,4,8}gpu_final.json and profiles/r02_bench_2gpu_a.json)
        wanted = [v for v in os.environ.get("B200_AQP_BENCH_VARIANTS", "").split(",") if v]
        for name, cls in variants.items():
            if cls is type(plan) or name not in wanted:
                continue
            other = cls(nR, nS, dev)
            for _ in range(3):
                sb = other.run(R, S)
            assert sb["matches"] == nS
            barrier()
            e0.record()
            xb = 0.0
            for _ in range(args.steps):
                sb = other.run(R, S)
                xb += sb["ms_exchange"] if name == "nccl" else sb["ms_pass1"]
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            exchange["variants"][name] = {"ms_per_step": float(t.item()), "value": (nR + nS) / float(t.item()) / 1e3,
                                          "unit": UNIT, "ms_exchange": xb / args.steps}
            if hasattr(other, "close"):
                other.close()
            del other

    # ---- e2e through run_join() with pinned host relations (N=1 path of the drop-in API) ----------
    e2e = None
    if world == 1:
        hR = torch.empty(nR * 2, dtype=torch.int32).pin_memory()
        hS = torch.empty(nS * 2, dtype=torch.int32).pin_memory()
        hR.copy_(R)
        hS.copy_(S)
        torch.cuda.synchronize()
        npR = hR.numpy().view(A.ROW).reshape(nR)
        npS = hS.numpy().view(A.ROW).reshape(nS)
        e2e_steps = max(1, min(args.steps, 5))
        for _ in range(2):
            g = A.run_join(npR, npS)
        assert g["matches"] == nS
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            g = A.run_join(npR, npS)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / e2e_steps
        e2e = {"value": (nR + nS) / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": 8 * (nR + nS),
               "d2h_bytes_per_step": 32, "ms_per_step": dt * 1e3, "steps": e2e_steps,
               "api": "run_join(result_t*, R, S, \"RHO\", joinconfig_t*) on pinned host relations "
                      "(what this library's create_relation_* return)"}
        # the other corners of (pinned | pageable inputs) x (count-only | materialised output), same call
        import numpy as np

        def wall(fn, reps):
            fn()
            t0 = time.perf_counter()
            for _ in range(reps):
                g = fn()
            return (time.perf_counter() - t0) / reps, g

        variants = {}
        # opt-in overlap (B200_AQP_E2E_CHUNKS): S travels in 8 chunks, R is joined with every chunk that has arrived
        os.environ["B200_AQP_E2E_CHUNKS"] = "8"
        dt2, g = wall(lambda: A.run_join(npR, npS), 3)
        assert g["matches"] == nS and g["checksum"] == s["checksum"]
        variants["pinned_count_overlapped"] = {"value": (nR + nS) / dt2 / 1e6, "unit": UNIT, "ms_per_step": dt2 * 1e3,
                                               "h2d_bytes_per_step": 8 * (nR + nS), "d2h_bytes_per_step": 32 * 8,
                                               "note": "B200_AQP_E2E_CHUNKS=8: copy of chunk k+1 under the join of R with chunk k "
                                                       "(R partitioned 8 times); not the default because run_join's reported "
                                                       "kernel throughput then includes the repeated R work"}
        pgR, pgS = np.empty(nR, dtype=A.ROW), np.empty(nS, dtype=A.ROW)   # plain malloc'd memory, as a reference caller has
        pgR[:] = npR
        pgS[:] = npS
        dt2, g = wall(lambda: A.run_join(pgR, pgS), 3)
        assert g["matches"] == nS
        variants["pageable_count_overlapped"] = {"value": (nR + nS) / dt2 / 1e6, "unit": UNIT, "ms_per_step": dt2 * 1e3,
                                                 "h2d_bytes_per_step": 8 * (nR + nS), "d2h_bytes_per_step": 32 * 8}
        del os.environ["B200_AQP_E2E_CHUNKS"]
        dt2, g = wall(lambda: A.run_join(pgR, pgS), 3)
        assert g["matches"] == nS
        variants["pageable_count"] = {"value": (nR + nS) / dt2 / 1e6, "unit": UNIT, "ms_per_step": dt2 * 1e3,
                                      "h2d_bytes_per_step": 8 * (nR + nS), "d2h_bytes_per_step": 32,
                                      "note": "multi-threaded staging through pinned buffers (csrc/hostcopy.cpp)"}
        del pgR, pgS
        # materialised output at BASELINE config 1 (2^24 x 2^26: 67 M triples = 805 MB of 16 KiB chunks handed back in
        # the reference's chunked_table_t layout); the reference's own numbers for this case: 740 / 929 Mtuples/s
        n1R, n1S = 1 << 24, 1 << 26
        d1R = torch.empty(n1R * 2, dtype=torch.int32, device=dev)
        d1S = torch.empty(n1S * 2, dtype=torch.int32, device=dev)
        A.gen_pk_device(d1R.data_ptr(), n1R, 11111, 0, n1R, st)
        A.gen_fk_device(d1S.data_ptr(), n1S, n1R, 22222, 0, n1S, st)
        torch.cuda.synchronize()
        for name, pin in (("pinned", True), ("pageable", False)):
            if pin:
                tR, tS = torch.empty(n1R * 2, dtype=torch.int32).pin_memory(), torch.empty(n1S * 2, dtype=torch.int32).pin_memory()
                tR.copy_(d1R)
                tS.copy_(d1S)
                aR, aS = tR.numpy().view(A.ROW).reshape(n1R), tS.numpy().view(A.ROW).reshape(n1S)
            else:
                aR, aS = np.empty(n1R, dtype=A.ROW), np.empty(n1S, dtype=A.ROW)
                aR[:] = d1R.cpu().numpy().view(A.ROW).reshape(n1R)
                aS[:] = d1S.cpu().numpy().view(A.ROW).reshape(n1S)
            torch.cuda.synchronize()
            for mat in (False, True):
                dt2, g = wall(lambda: A.run_join(aR, aS, materialize=mat, keep_triples=False), 3)
                assert g["matches"] == n1S and (not mat or g["table_num_tuples"] == n1S)
                variants[f"c1_{name}_{'materialised' if mat else 'count'}"] = {
                    "value": (n1R + n1S) / dt2 / 1e6, "unit": UNIT, "ms_per_step": dt2 * 1e3,
                    "h2d_bytes_per_step": 8 * (n1R + n1S), "d2h_bytes_per_step": 12 * n1S + 8 * (n1S // 1364 + 1) if mat else 32}
            del aR, aS
        del d1R, d1S
        e2e["variants"] = variants
        e2e["variants_note"] = ("c1_* = BASELINE config 1 (2^24 x 2^26) through run_join(); materialised = MATERIALIZE=1, the "
                                "result handed back as the reference's chunked_table_t (16 KiB chunks in one host slab) and "
                                "released with destroy_table()")
        del hR, hS, npR, npS
    else:
        # N > 1: every rank keeps its row-range shard in pinned host memory; a step = H2D of the shard over this
        # GPU's own PCIe link + the sharded join (whose result read-back is the final 24-byte all-reduce)
        hR = torch.empty(nR_loc * 2, dtype=torch.int32).pin_memory()
        hS = torch.empty(nS_loc * 2, dtype=torch.int32).pin_memory()
        hR.copy_(R)
        hS.copy_(S)
        torch.cuda.synchronize()
        e2e_steps = max(1, min(args.steps, 5))

        def e2e_step():
            R.copy_(hR, non_blocking=True)
            S.copy_(hS, non_blocking=True)
            return plan.run(R, S)

        for _ in range(2):
            g = e2e_step()
        assert g["matches"] == nS
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            g = e2e_step()
        barrier()
        t = torch.tensor([(time.perf_counter() - t0) / e2e_steps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
        e2e = {"value": (nR + nS) / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": 8 * (nR + nS),
               "d2h_bytes_per_step": 24 * world, "ms_per_step": dt * 1e3, "steps": e2e_steps,
               "api": f"{type(plan).__name__}.run on pinned host row-range shards, one H2D stream per GPU"}
        del hR, hS
    del R, S
    torch.cuda.empty_cache()

    scan = bench_scan(A, torch, dev, peak, max(args.steps, 5), args.warmup, world, rank)
    if world == 1:
        import numpy as np
        scan["e2e"] = bench_scan_e2e(A, np, 3)
    if world > 1:
        for k, v in scan.items():
            if isinstance(v, dict) and "ms" in v:
                t = torch.tensor([v["ms"]], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                r = v["roofline"]
                r["achieved"] *= v["ms"] / float(t.item())   # per GPU, at the slowest rank's time
                r["frac"] = r["achieved"] / r["peak"]
                v["ms"] = float(t.item())
                n_tot = 1 << SCAN_LOG_N
                v["input_gbs"] = n_tot / v["ms"] / 1e6

    # ---- BASELINE config 5: TPC-H-style Q3 / Q12 / Q19 at SF100 on device-generated columns (1 GPU) -------
    tpch = None
    if world == 1 and not args.no_tpch:
        sf = float(os.environ.get("B200_AQP_TPCH_SF", "100"))
        A.tpch_generate_device(sf, 1)
        published_ms = {3: 730.0, 12: 225.0, 19: 120.0}   # reference, 16 threads, SF100, other hardware
        tpch = {"scale_factor": sf, "data": "synthetic, generated in HBM (include/aqp/b200_tpch.h)",
                "roofline_note": "algorithmic_bytes counts every column a selection references once; a selection that rejects "
                                 "on a one-byte code never touches most sectors of its wide columns, so frac can exceed 1 (Q19)"}
        nc, no, npart = int(150000 * sf), int(1500000 * sf), int(200000 * sf)
        nl = 4 * no

        def tpch_bytes(q, r):
            """Algorithmic HBM bytes of a pipeline: every column a filter has to read once (key+row-id pairs 8 B, dates
            8 B, codes 1 B, part key / size / quantity 4 B), 8 B per selected row written, the joins at the graded
            56 B/tuple, 12 B per materialised triple written and read again by its consumer."""
            f, j1 = r["filtered"], r["join1_rows"]
            join = lambda a, b: JOIN_BYTES_PER_TUPLE * (a + b)
            if q == 12:   # Q12Predicates.hpp:40-109: l_orderkey, l_shipmode, 3 dates
                return nl * (8 + 1 + 24) + 8 * f[0] + join(no, f[0])
            if q == 3:    # Q3Predicates.hpp:57-184
                return (nc * (8 + 1) + 8 * f[0] + no * (8 + 8 + 4) + 8 * f[1] + join(f[0], f[1]) + 12 * j1 +
                        (12 + 8) * j1 + nl * (8 + 8) + 8 * f[2] + join(j1, f[2]))
            # Q19Predicates.hpp:58-78,:194-389; the post-join check gathers 10 B of columns per match by row id
            return npart * (8 + 1 + 1 + 4) + 8 * f[0] + nl * (8 + 4 + 4 + 1 + 1) + 8 * f[1] + join(f[0], f[1]) + 12 * j1 + (12 + 10) * j1

        for q in (3, 12, 19):
            for _ in range(2):
                r = A.tpch_query_device(q)
            runs = [A.tpch_query_device(q) for _ in range(3)]
            ms = sum(x["ms_total"] for x in runs) / len(runs)
            b = tpch_bytes(q, r)
            tpch[f"q{q}"] = {"ms": ms, "mrows_per_s": r["input_rows"] / ms / 1e3, "result_rows": r["result_rows"],
                             "ms_filter": runs[-1]["ms_filter"], "ms_join": runs[-1]["ms_join"],
                             "filtered": r["filtered"], "join1_rows": r["join1_rows"],
                             "roofline": {"bound": "hbm", "algorithmic_bytes": b, "achieved": b / ms / 1e6, "peak": peak,
                                          "unit": "GB/s", "frac": b / ms / 1e6 / peak},
                             "reference_published_ms_sf100": published_ms[q]}
        if not args.no_cpu_baseline:
            try:   # the reference's own pipelines on the host cores, bounded sample: the same generator at SF1
                import oracle as O
                if O.have_ref():
                    sf_cpu = float(os.environ.get("B200_AQP_TPCH_CPU_SF", "10"))
                    A.tpch_generate_device(sf_cpu, 1)
                    t = A.tpch_download()
                    tpch["cpu_reference"] = {"cores": os.cpu_count(), "kind": "reference", "scale_factor": sf_cpu,
                                                 "note": "the reference's unmodified pipelines (oracle/_ref/libref_tpch.so) on the "
                                                         "downloaded tables; result rows asserted equal to the device pipelines'"}
                    for q in (3, 12, 19):
                        g = A.tpch_query_device(q)
                        c = O.ref_tpch_query(q, t, nthreads=os.cpu_count())
                        assert c["result_rows"] == g["result_rows"], (q, c, g)
                        tpch["cpu_reference"][f"q{q}"] = {"ms": c["seconds"] * 1e3, "gpu_ms": g["ms_total"],
                                                              "result_rows": c["result_rows"]}
                    del t
            except Exception as ex:
                tpch["cpu_reference"] = {"failed": str(ex)}
        A.lib().b200_tpch_free_device()

    # ---- the same three pipelines sharded over the N GPUs: every rank generates and filters rows
    # [total * rank / N, total * (rank + 1) / N) of each table, the joins are the sharded join of csrc/mg.cu (Q3: the
    # matches of join 1 stay sharded by key and build join 2; Q19: the final predicate's attributes travel in the payloads)
    if world > 1 and not args.no_tpch:
        if hasattr(plan, "close"):
            plan.close()   # one multi-GPU host per process: the join's goes, TPC-H's comes
        plan = None
        sf = float(os.environ.get("B200_AQP_TPCH_SF", "100"))
        A.tpch_generate_shard_device(sf, 1, rank, world)
        uid = torch.zeros(128, dtype=torch.uint8, device=dev)
        if rank == 0:
            uid.copy_(torch.tensor(list(A.mg_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        A.tpch_mg_init(rank, world, bytes(uid.cpu().tolist()))
        fns = {3: A.tpch_q3_mg, 12: A.tpch_q12_mg, 19: A.tpch_q19_mg}
        tpch = {"scale_factor": sf, "sharding": f"row ranges of every table over {world} GPUs, generated in HBM",
                "data": "synthetic (include/aqp/b200_tpch.h)"}
        for q in (3, 12, 19):
            for _ in range(2):
                r = fns[q]()
            runs = [fns[q]() for _ in range(3)]
            t = torch.tensor([sum(x["ms_total"] for x in runs) / len(runs)], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            rows_in = torch.tensor([r["input_rows"]], dtype=torch.int64, device=dev)
            dist.all_reduce(rows_in)
            ms = float(t.item())
            tpch[f"q{q}"] = {"ms": ms, "mrows_per_s": int(rows_in.item()) / ms / 1e3, "result_rows": r["result_rows"],
                             "join1_rows": r["join1_rows"], "ms_filter_rank0": runs[-1]["ms_filter"],
                             "ms_join_rank0": runs[-1]["ms_join"]}
        A.mg_finalize()
        if rank == 0:   # the single-GPU pipelines on the whole data set give the answers the sharded ones must match
            A.tpch_generate_device(sf, 1)
            for q in (3, 12, 19):
                g = A.tpch_query_device(q)
                assert (g["result_rows"], g["join1_rows"]) == (tpch[f"q{q}"]["result_rows"], tpch[f"q{q}"]["join1_rows"]), (q, g, tpch)
                tpch[f"q{q}"]["one_gpu_ms_cold"] = g["ms_total"]   # first call, no warm-up: a check, not a measurement
            tpch["check"] = "result and join-1 rows equal to the single-GPU pipelines' on the full tables (run on rank 0)"
        A.lib().b200_tpch_free_device()
        dist.barrier()

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                cpu = cpu_reference_join(3, 1)
                cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "wall_ms_per_step", "tsc_hz",
                                           "same_size_as_headline", "host")}
                cs = cpu_reference_scan()
                if cs:
                    cpu["scan"] = cs
            except Exception as ex:   # the baseline is reporting only; never fail the bench on it
                cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {ex}"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "u32", "data": "synthetic",
                "config": workload_config(world),
                "plan": {"generator": "on-device bijection, seeds 11111/22222", "radix_bits": s["radix_bits"],
                         "passes": s["num_passes"], "histogram_free": bool(s.get("plan_flags", 0) & 1),
                         "exchange": None if world == 1 else
                         f"pass 1 routes by the low key bits and stores every run into the owner's buffer over NVLink peer "
                         f"memory ({s.get('exchange', 'nccl')}); NCCL carries the sizing collectives"},
                "phases_ms": phase, "roofline": roof, "join_roofline": join_roof, "cpu_baseline": cpu, "e2e": e2e,
                "exchange": exchange, "skew": skew, "tpch": tpch,
                "gpu_launches": launches_timed, "gpu_launches_total": A.kernel_launch_count() - launches0,
                "clocks": clocks, "scan": scan}
        print(json.dumps(line))
    if world > 1:
        if plan is not None and hasattr(plan, "close"):
            plan.close()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tpch", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
