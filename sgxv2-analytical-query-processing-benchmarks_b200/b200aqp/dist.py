"""Sharded RHO join across the GPUs of one box: one process per GPU, torch.distributed for the plumbing.

The reference is a single shared-memory process (no communication backend); its
inter-thread "shuffle" is the pass-1 scatter into one shared array (radix_join.cpp:901-926). Here the
same pass is the inter-GPU shuffle:

  1. every rank holds a row range of R and S;  b200_shard_pass1_device histograms it and scatters it
     by the ROUTED pass-1 digit, so the partitions owned by GPU g (low log2(G) key bits == g) form one
     contiguous range of the output — the output IS the all-to-all send buffer;
  2. two small collectives size the exchange: an all-gather of the pass-1 partition counts (who
     sends how much of which partition to whom) and an all-reduce of the full-width histograms (final
     partition sizes);
  3. the 8-byte tuples move with one all-to-all per relation (NCCL over NVLink 5 / NVSwitch);
  4. every rank finishes alone: b200_shard_join_device runs pass 2 over the received segments and the
     shared-memory build/probe;  5. a 24-byte all-reduce sums matches / checksum / keysum.

All device work goes through the C ABI (`backend`); this module only computes split sizes, segment
tables and histogram slices (pure torch, device-agnostic — tests/test_dist_gloo.py runs it with gloo
and a host stand-in for the kernels).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def log2_exact(x: int) -> int:
    l = x.bit_length() - 1
    if x <= 0 or (1 << l) != x:
        raise ValueError(f"{x} is not a power of two")
    return l


def plan_bits(nR_total: int, world: int, join_plan) -> tuple[int, int, int]:
    """Radix bits of the sharded join: the single-GPU plan for the WHOLE build side, widened so that
    pass 1 has at least log2(world) bits to route on."""
    bits, b1, b2 = join_plan(nR_total)
    lg = log2_exact(world)
    if b1 < lg:
        b1 = lg
        bits = max(bits, lg)
        b2 = bits - b1
    return bits, b1, b2


def exchange_plan(counts_all: torch.Tensor, rank: int, world: int):
    """counts_all[s, p] = tuples rank s holds of routed pass-1 partition p (all-gathered).
    Returns (send_splits, recv_splits, seg_off, seg_group):
      send_splits[g]  tuples this rank sends to g        = its partitions in g's range
      recv_splits[s]  tuples this rank receives from s   = s's partitions in this rank's range
      seg_off         [world * per + 1] starts of the received segments, ordered (source, partition)
      seg_group       [world * per]     local partition index of every segment"""
    F1 = counts_all.shape[1]
    per = F1 // world
    mine = counts_all[rank].view(world, per).sum(1)
    incoming = counts_all[:, rank * per:(rank + 1) * per]            # [world, per]
    recv = incoming.sum(1)
    seg_len = incoming.reshape(-1)
    seg_off = torch.zeros(seg_len.numel() + 1, dtype=torch.int64, device=counts_all.device)
    seg_off[1:] = torch.cumsum(seg_len, 0)
    seg_group = torch.arange(per, device=counts_all.device, dtype=torch.int32).repeat(world)
    return mine, recv, seg_off, seg_group


def dest_offsets(counts_all: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """Fused exchange: dest_off[p] = tuple index inside the receive buffer of partition p's owner at which
    THIS rank's segment of p starts. Every owner lays its buffer out by (source rank, local partition),
    the same order exchange_plan() reports as seg_off on the receiving side."""
    F1 = counts_all.shape[1]
    per = F1 // world
    c = counts_all.view(world, world, per)                       # [source, owner, local partition]
    per_src_owner = c.sum(2)                                     # tuples source s sends to owner g
    before = torch.cumsum(per_src_owner, 0) - per_src_owner      # sent to g by lower-ranked sources
    mine = c[rank]                                               # [owner, per]
    within = torch.cumsum(mine, 1) - mine
    return (before[rank].unsqueeze(1) + within).reshape(-1)


def final_hist_slice(hist_global: torch.Tensor, rank: int, world: int, b1: int, b2: int) -> torch.Tensor:
    """hist_global is indexed by the routed full-width digit p1' | (p2 << b1). Returns this rank's slice
    in final partition order (local pass-1 partition major, then p2): [per << b2]."""
    F1, F2 = 1 << b1, 1 << b2
    per = F1 // world
    return hist_global.view(F2, F1)[:, rank * per:(rank + 1) * per].t().contiguous().view(-1)


class CudaBackend:
    """The product backend: libb200aqp.so through ctypes."""

    def __init__(self):
        import b200aqp as A
        self.A = A
        self.L = A.lib()

    def stream(self):
        return self.A._st(torch.cuda.current_stream().cuda_stream)

    def shard_pass1(self, rel, n, bits, b1, lg, send, hist, part1_off):
        self.A._check(self.L.b200_shard_pass1_device(rel.data_ptr(), n, bits, b1, lg, send.data_ptr(), hist.data_ptr(),
                                                     part1_off.data_ptr(), self.stream()), "b200_shard_pass1_device")

    def shard_join(self, R, nR, segR, S, nS, segS, seg_group, nseg, ngroups, shift2, bits2, histR, histS, hash_shift):
        import ctypes as C
        s = self.A.JoinStats()
        self.A._check(self.L.b200_shard_join_device(R.data_ptr(), nR, segR.data_ptr(), S.data_ptr(), nS, segS.data_ptr(),
                                                    seg_group.data_ptr(), nseg, ngroups, shift2, bits2, histR.data_ptr(),
                                                    histS.data_ptr(), hash_shift, C.byref(s), self.stream()),
                      "b200_shard_join_device")
        return s.as_dict()

    def join_plan(self, nR):
        return self.A.join_plan(nR)

    # ---- fused scatter + exchange over peer memory ---------------------------------------------------
    def shard_hist(self, rel, n, bits, b1, lg, hist, counts1, slot):
        self.A._check(self.L.b200_shard_hist_device(rel.data_ptr(), n, bits, b1, lg, hist.data_ptr(), counts1.data_ptr(),
                                                    slot, self.stream()), "b200_shard_hist_device")

    def shard_scatter(self, rel, n, dest_off, dest_ptrs, slot):
        import ctypes as C
        arr = (C.c_void_p * len(dest_ptrs))(*dest_ptrs)
        self.A._check(self.L.b200_shard_scatter_device(rel.data_ptr(), n, dest_off.data_ptr(), arr, slot, self.stream()),
                      "b200_shard_scatter_device")

    def exchange_plan(self, counts_all, world, rank, b1, b2, hist_global, seg_off, dest_off, hist_slice, host_vals):
        self.A._check(self.L.b200_exchange_plan_device(counts_all.data_ptr(), world, rank, b1, b2, hist_global.data_ptr(),
                                                       seg_off.data_ptr(), dest_off.data_ptr(), hist_slice.data_ptr(),
                                                       host_vals.data_ptr(), self.stream()), "b200_exchange_plan_device")

    def shard_join_async(self, R, nR, segR, S, nS, segS, seg_group, nseg, ngroups, shift2, bits2, histR, histS, hash_shift,
                         result3):
        self.A._check(self.L.b200_shard_join_async_device(R.data_ptr(), nR, segR.data_ptr(), S.data_ptr(), nS,
                                                          segS.data_ptr(), seg_group.data_ptr(), nseg, ngroups, shift2,
                                                          bits2, histR.data_ptr(), histS.data_ptr(), hash_shift,
                                                          result3.data_ptr(), self.stream()),
                      "b200_shard_join_async_device")

    def shard_join_times(self):
        s = self.A.JoinStats()
        self.A._check(self.L.b200_shard_join_times(s), "b200_shard_join_times")
        return s.as_dict()

    def copy_async(self, dst_ptr, src_ptr, nbytes, stream):
        self.A._check(self.L.b200_copy_async(dst_ptr, src_ptr, nbytes, self.A._st(stream)), "b200_copy_async")

    def alloc_shared(self, nbytes):
        """(buffer, 64-byte IPC handle as a uint8 tensor)"""
        import ctypes as C
        buf = self.A.DeviceBuffer(nbytes)
        h = (C.c_ubyte * 64)()
        self.A._check(self.L.b200_ipc_export(buf.ptr, h), "b200_ipc_export")
        return buf, torch.tensor(list(h), dtype=torch.uint8)

    def close_shared(self, ptr: int):
        self.A._check(self.L.b200_ipc_close(ptr), "b200_ipc_close")

    def open_shared(self, handle: torch.Tensor) -> int:
        import ctypes as C
        h = (C.c_ubyte * 64)(*handle.tolist())
        p = C.c_void_p()
        self.A._check(self.L.b200_ipc_open(h, C.byref(p)), "b200_ipc_open")
        return p.value


class ShardedJoin:
    """R and S are int32 tensors of shape [2 * n_local] (key, payload interleaved = row_t) holding this
    rank's row range. run() returns the GLOBAL matches / checksum / keysum plus per-phase device times."""

    def __init__(self, nR_total: int, nS_total: int, device, backend=None, group=None):
        self.backend = backend or CudaBackend()
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = device
        self.nR_total, self.nS_total = nR_total, nS_total
        self.lg = log2_exact(self.world)
        self.bits, self.b1, self.b2 = plan_bits(nR_total, self.world, self.backend.join_plan)
        self.F1, self.P = 1 << self.b1, 1 << self.bits
        self._bufs = {}

    def _buf(self, name, numel, dtype):
        t = self._bufs.get(name)
        if t is None or t.numel() < numel or t.dtype != dtype:
            t = torch.empty(max(numel, 2), dtype=dtype, device=self.device)
            self._bufs[name] = t
        return t

    def _event(self):
        if self.device.type != "cuda":
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def run(self, R: torch.Tensor, S: torch.Tensor) -> dict:
        be, G, rank = self.backend, self.world, self.rank
        nR, nS = R.numel() // 2, S.numel() // 2
        F1, P = self.F1, self.P
        e0 = self._event()
        # ---- 1. local histogram + routed pass-1 scatter (output = send buffer) -----------------------
        sendR = self._buf("sendR", 2 * nR + 4, torch.int32)
        sendS = self._buf("sendS", 2 * nS + 4, torch.int32)
        hist = self._buf("hist", 2 * P, torch.int32)
        off1 = self._buf("off1", 2 * (F1 + 1), torch.int32)
        histR, histS = hist[:P], hist[P:2 * P]
        offR, offS = off1[:F1 + 1], off1[F1 + 1:2 * (F1 + 1)]
        be.shard_pass1(R, nR, self.bits, self.b1, self.lg, sendR, histR, offR)
        be.shard_pass1(S, nS, self.bits, self.b1, self.lg, sendS, histS, offS)
        e1 = self._event()
        # ---- 2. size the exchange ---------------------------------------------------------------------
        counts = torch.cat([offR[1:] - offR[:-1], offS[1:] - offS[:-1]]).to(torch.int64)       # [2*F1]
        counts_flat = torch.empty(G * 2 * F1, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(counts_flat, counts, group=self.group)
        counts_all = counts_flat.view(G, 2 * F1)
        hist_global = hist[:2 * P].clone()
        dist.all_reduce(hist_global, group=self.group)
        sR, rR, segR, seg_group = exchange_plan(counts_all[:, :F1], rank, G)
        sS, rS, segS, _ = exchange_plan(counts_all[:, F1:], rank, G)
        splits = torch.stack([sR, rR, sS, rS]).tolist()        # the one host sync of the exchange
        nR_recv, nS_recv = int(sum(splits[1])), int(sum(splits[3]))
        # ---- 3. all-to-all of the 8-byte tuples ---------------------------------------------------------
        recvR = self._buf("recvR", 2 * nR_recv + 4, torch.int32)
        recvS = self._buf("recvS", 2 * nS_recv + 4, torch.int32)
        dist.all_to_all_single(recvR[:2 * nR_recv].view(torch.int64), sendR[:2 * nR].view(torch.int64),
                               output_split_sizes=splits[1], input_split_sizes=splits[0], group=self.group)
        dist.all_to_all_single(recvS[:2 * nS_recv].view(torch.int64), sendS[:2 * nS].view(torch.int64),
                               output_split_sizes=splits[3], input_split_sizes=splits[2], group=self.group)
        e2 = self._event()
        # ---- 4. local pass 2 + build/probe ------------------------------------------------------------------
        per = F1 // G
        hR = final_hist_slice(hist_global[:P], rank, G, self.b1, self.b2)
        hS = final_hist_slice(hist_global[P:], rank, G, self.b1, self.b2)
        local = be.shard_join(recvR, nR_recv, segR.to(torch.int32), recvS, nS_recv, segS.to(torch.int32), seg_group,
                              G * per, per, self.b1, self.b2, hR, hS, self.bits)
        e3 = self._event()
        # ---- 5. global result ---------------------------------------------------------------------------------
        to_i64 = lambda v: v - (1 << 64) if v >= (1 << 63) else v
        res = torch.tensor([local["matches"], to_i64(local["checksum"]), to_i64(local["keysum"])], dtype=torch.int64,
                           device=self.device)
        dist.all_reduce(res, group=self.group)
        m, cs, ks = (int(x) for x in res.tolist())
        out = {"matches": m, "checksum": cs % (1 << 64), "keysum": ks % (1 << 64), "radix_bits": self.bits,
               "num_passes": 2, "bits_pass1": self.b1, "bits_pass2": self.b2, "tuples_sent": int(sum(splits[0]) + sum(splits[2])),
               "tuples_kept": int(splits[0][rank] + splits[2][rank]), "ms_pass2": local.get("ms_pass2", 0.0),
               "ms_join": local.get("ms_join", 0.0)}
        if e0 is not None:
            torch.cuda.synchronize()
            out["ms_pass1"] = e0.elapsed_time(e1)     # histogram + pass-1 scatter
            out["ms_hist"] = 0.0
            out["ms_exchange"] = e1.elapsed_time(e2)   # count/histogram collectives + all-to-all
            out["ms_total"] = e0.elapsed_time(e3)
        return out


class _RawBuffer:
    """Pointer + size with the .data_ptr() face the backends expect (memory owned by the C library)."""

    def __init__(self, ptr):
        self.ptr = ptr

    def data_ptr(self):
        return self.ptr


class FusedShardedJoin(ShardedJoin):
    """Same join, but the shuffle is fused into the pass-1 scatter kernel: every rank maps its peers'
    receive buffers (CUDA IPC over NVLink 5 / NVSwitch peer memory) and the kernel stores each tile's runs
    straight into the owner's buffer. No send buffer, no NCCL all-to-all; NCCL only carries the two small
    sizing collectives, one barrier and the final 24-byte all-reduce.

    capacity_factor sizes the receive buffers relative to a perfectly even split; if a run would
    overflow them (heavy skew) that run falls back to the NCCL path of the base class."""

    def __init__(self, nR_total, nS_total, device, backend=None, group=None, capacity_factor: float = 1.5):
        super().__init__(nR_total, nS_total, device, backend, group)
        be, G = self.backend, self.world
        self.capR = int(nR_total / G * capacity_factor) + 4096
        self.capS = int(nS_total / G * capacity_factor) + 4096
        self.peerR, self.peerS = [], []
        self._own = []
        for cap, peers in ((self.capR, self.peerR), (self.capS, self.peerS)):
            buf, handle = be.alloc_shared(cap * 8 + 64)
            self._own.append(buf)
            hs = torch.empty(G * 64, dtype=torch.uint8, device=device)
            dist.all_gather_into_tensor(hs, handle.to(device), group=self.group)
            hs = hs.cpu().view(G, 64)
            for g in range(G):
                peers.append(buf.ptr if g == self.rank else be.open_shared(hs[g]))
        self.fallbacks = 0
        self._seg_group = None

    def close(self):
        """Collective: unmap the peers' buffers, then free this rank's own (nobody may still have it mapped)."""
        if not self._own:
            return
        if self.device.type == "cuda":
            torch.cuda.synchronize()
        dist.barrier(group=self.group)            # every rank is done reading and writing peer memory
        if hasattr(self.backend, "close_shared"):
            for peers in (self.peerR, self.peerS):
                for g, ptr in enumerate(peers):
                    if g != self.rank:
                        self.backend.close_shared(ptr)
        dist.barrier(group=self.group)            # every mapping of this rank's buffers is closed
        self._own = []
        self.peerR, self.peerS = [], []

    def _sizing(self, cnt1, hist):
        """The two sizing collectives and everything derived from them. Returns (host, segR, segS, dR, dS, hR, hS,
        seg_group) with host = [largest receive size anywhere R, S; my receive sizes R, S; tuples I keep R, S] — reading
        it is the one host sync of the exchange. On the CUDA backend the derivation is ONE kernel
        (b200_exchange_plan_device); other backends (CPU tests) use the tensor formulation above."""
        be, G, rank, F1, P = self.backend, self.world, self.rank, self.F1, self.P
        per = F1 // G
        fast = hasattr(be, "exchange_plan")
        counts_flat = self._buf("counts_all", G * 2 * F1, torch.int32 if fast else torch.int64)
        dist.all_gather_into_tensor(counts_flat[:G * 2 * F1], cnt1[:2 * F1] if fast else cnt1[:2 * F1].to(torch.int64),
                                    group=self.group)
        dist.all_reduce(hist[:2 * P], group=self.group)          # in place: shard_hist re-zeroes it next run
        if self._seg_group is None:
            self._seg_group = torch.arange(per, device=self.device, dtype=torch.int32).repeat(G)
        if fast:
            nseg = G * per
            seg = self._buf("seg_off", 2 * (nseg + 1), torch.int32)
            dest = self._buf("dest_off", 2 * F1, torch.int32)
            hsl = self._buf("hist_slice", 2 * (per << self.b2), torch.int32)
            hv = self._buf("host_vals", 6, torch.int64)
            be.exchange_plan(counts_flat, G, rank, self.b1, self.b2, hist, seg, dest, hsl, hv)
            host = hv[:6].tolist()
            n2 = per << self.b2
            return (host, seg[:nseg + 1], seg[nseg + 1:2 * (nseg + 1)], dest[:F1], dest[F1:2 * F1], hsl[:n2],
                    hsl[n2:2 * n2], self._seg_group)
        counts_all = counts_flat[:G * 2 * F1].view(G, 2 * F1)
        _, rR, segR, _ = exchange_plan(counts_all[:, :F1], rank, G)
        _, rS, segS, _ = exchange_plan(counts_all[:, F1:], rank, G)
        dR = dest_offsets(counts_all[:, :F1], rank, G).to(torch.int32)
        dS = dest_offsets(counts_all[:, F1:], rank, G).to(torch.int32)
        # every rank must take the same path: the largest receive sizes anywhere are compared with the capacities
        host = torch.stack([counts_all[:, :F1].view(G, G, per).sum((0, 2)).max(),
                            counts_all[:, F1:].view(G, G, per).sum((0, 2)).max(), rR.sum(), rS.sum(),
                            counts_all[rank, rank * per:(rank + 1) * per].sum(),
                            counts_all[rank, F1 + rank * per:F1 + (rank + 1) * per].sum()]).tolist()
        hR = final_hist_slice(hist[:P], rank, G, self.b1, self.b2)
        hS = final_hist_slice(hist[P:2 * P], rank, G, self.b1, self.b2)
        return host, segR.to(torch.int32), segS.to(torch.int32), dR, dS, hR, hS, self._seg_group

    def _finish(self, nR_recv, segR, nS_recv, segS, seg_group, hR, hS):
        """Local pass 2 + build/probe on the received segments and the global sum of the result.
        Returns (matches, checksum, keysum, phase times of the local stage)."""
        be, G, rank = self.backend, self.world, self.rank
        per = self.F1 // G
        bufR, bufS = _RawBuffer(self.peerR[rank]), _RawBuffer(self.peerS[rank])
        if hasattr(be, "shard_join_async"):
            res = self._buf("result3", 4, torch.int64)
            be.shard_join_async(bufR, nR_recv, segR, bufS, nS_recv, segS, seg_group, G * per, per, self.b1, self.b2,
                                hR, hS, self.bits, res)
            dist.all_reduce(res[:3], group=self.group)
            m, cs, ks = (int(x) for x in res[:3].tolist())       # the sync that ends the join
            local = be.shard_join_times()
        else:
            local = be.shard_join(bufR, nR_recv, segR, bufS, nS_recv, segS, seg_group, G * per, per, self.b1, self.b2,
                                  hR, hS, self.bits)
            to_i64 = lambda v: v - (1 << 64) if v >= (1 << 63) else v
            res = torch.tensor([local["matches"], to_i64(local["checksum"]), to_i64(local["keysum"])], dtype=torch.int64,
                               device=self.device)
            dist.all_reduce(res, group=self.group)
            m, cs, ks = (int(x) for x in res.tolist())
        return m, cs % (1 << 64), ks % (1 << 64), local

    def run(self, R: torch.Tensor, S: torch.Tensor) -> dict:
        be, G, rank = self.backend, self.world, self.rank
        nR, nS = R.numel() // 2, S.numel() // 2
        F1, P = self.F1, self.P
        e0 = self._event()
        # ---- 1. local histograms (per-CTA rows stay inside the library for the scatter) --------------
        hist = self._buf("hist", 2 * P, torch.int32)
        cnt1 = self._buf("cnt1", 2 * F1, torch.int32)
        be.shard_hist(R, nR, self.bits, self.b1, self.lg, hist[:P], cnt1[:F1], 0)
        be.shard_hist(S, nS, self.bits, self.b1, self.lg, hist[P:2 * P], cnt1[F1:2 * F1], 1)
        eh = self._event()
        # ---- 2. size the exchange -------------------------------------------------------------------------
        host, segR, segS, dR, dS, hR, hS, seg_group = self._sizing(cnt1, hist)
        if host[0] > self.capR or host[1] > self.capS:
            self.fallbacks += 1
            out = ShardedJoin.run(self, R, S)
            out["exchange"] = "nccl-fallback"
            return out
        nR_recv, nS_recv = int(host[2]), int(host[3])
        es = self._event()
        # ---- 3. fused scatter + exchange: stores go to the owners' buffers over NVLink ------------------
        be.shard_scatter(R, nR, dR, self.peerR, 0)
        be.shard_scatter(S, nS, dS, self.peerS, 1)
        ex = self._event()
        flag = self._buf("flag", 2, torch.int32)
        dist.all_reduce(flag[:1], group=self.group)    # barrier: every rank's stores have landed
        e2 = self._event()
        # ---- 4. local pass 2 + build/probe, global result ----------------------------------------------------
        m, cs, ks, local = self._finish(nR_recv, segR, nS_recv, segS, seg_group, hR, hS)
        e3 = self._event()
        out = {"matches": m, "checksum": cs, "keysum": ks, "radix_bits": self.bits,
               "num_passes": 2, "bits_pass1": self.b1, "bits_pass2": self.b2, "tuples_sent": nR + nS,
               "tuples_kept": int(host[4] + host[5]), "ms_pass2": local.get("ms_pass2", 0.0),
               "ms_join": local.get("ms_join", 0.0), "exchange": "p2p-fused"}
        if e0 is not None:
            torch.cuda.synchronize()
            out["ms_hist"] = e0.elapsed_time(eh)
            out["ms_pass1"] = eh.elapsed_time(e2)      # sizing collectives + fused scatter/exchange + barrier
            out["ms_sizing"] = eh.elapsed_time(es)     # all-gather + all-reduce + plan kernel + host sync
            out["ms_scatter_kernels"] = es.elapsed_time(ex)
            out["ms_barrier"] = ex.elapsed_time(e2)
            out["ms_exchange"] = 0.0                   # no separate exchange step
            out["ms_total"] = e0.elapsed_time(e3)      # up to the global result
        return out


class DmaShardedJoin(FusedShardedJoin):
    """Third exchange variant: pass 1 scatters into a LOCAL send buffer laid out by destination, then one
    copy-engine transfer per (relation, destination) moves each block straight into the peer's IPC-mapped
    receive buffer. Copy engines move large contiguous blocks at the full NVLink rate (774 GB/s measured peer
    copy, profiles/r01_p2pbench_2gpu.txt) where SM-issued 256-byte runs reach 300-435 GB/s; the price is the
    send buffer's extra HBM write + read. Overlap: the sizing collectives and the plan arithmetic run on a side
    stream under the pass-1 scatter (the send-buffer layout needs only local counts), and R's transfers run
    under S's scatter."""

    def __init__(self, nR_total, nS_total, device, backend=None, group=None, capacity_factor: float = 1.5):
        super().__init__(nR_total, nS_total, device, backend, group, capacity_factor)
        cuda = device.type == "cuda"
        self.side = torch.cuda.Stream(device=device) if cuda else None
        self.copy_streams = [torch.cuda.Stream(device=device) for _ in range(2 * self.world)] if cuda else []

    def _plan(self, cnt1, hist):
        """sizing collectives + plan arithmetic (device tensors only); returns the plan and ONE flat int64
        tensor with everything the host needs to issue the transfers"""
        G, rank, F1, P = self.world, self.rank, self.F1, self.P
        per = F1 // G
        counts_flat = torch.empty(G * 2 * F1, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(counts_flat, cnt1[:2 * F1].to(torch.int64), group=self.group)
        counts_all = counts_flat.view(G, 2 * F1)
        hist_global = hist[:2 * P].clone()
        dist.all_reduce(hist_global, group=self.group)
        sR, rR, segR, seg_group = exchange_plan(counts_all[:, :F1], rank, G)
        sS, rS, segS, _ = exchange_plan(counts_all[:, F1:], rank, G)
        # where my block for destination g starts inside g's receive buffer = tuples sent to g by lower ranks
        toR = counts_all[:, :F1].view(G, G, per).sum(2)          # [source, owner]
        toS = counts_all[:, F1:].view(G, G, per).sum(2)
        dstR = (torch.cumsum(toR, 0) - toR)[rank]
        dstS = (torch.cumsum(toS, 0) - toS)[rank]
        worst = torch.stack([toR.sum(0).max(), toS.sum(0).max(), rR.sum(), rS.sum()])
        hv = torch.cat([worst, sR, sS, dstR, dstS])
        hR = final_hist_slice(hist_global[:P], rank, G, self.b1, self.b2)
        hS = final_hist_slice(hist_global[P:], rank, G, self.b1, self.b2)
        return hv, segR.to(torch.int32), segS.to(torch.int32), seg_group, hR, hS

    def run(self, R: torch.Tensor, S: torch.Tensor) -> dict:
        be, G, rank = self.backend, self.world, self.rank
        nR, nS = R.numel() // 2, S.numel() // 2
        F1, P = self.F1, self.P
        per = F1 // G
        cuda = self.device.type == "cuda"
        main = torch.cuda.current_stream() if cuda else None
        e0 = self._event()
        sendR = self._buf("sendR", 2 * nR + 4, torch.int32)
        sendS = self._buf("sendS", 2 * nS + 4, torch.int32)
        hist = self._buf("hist", 2 * P, torch.int32)
        cnt1 = self._buf("cnt1", 2 * F1, torch.int32)
        # ---- 1. local histograms --------------------------------------------------------------------------
        be.shard_hist(R, nR, self.bits, self.b1, self.lg, hist[:P], cnt1[:F1], 0)
        be.shard_hist(S, nS, self.bits, self.b1, self.lg, hist[P:2 * P], cnt1[F1:2 * F1], 1)
        eh = self._event()
        # ---- 2. sizing on the side stream, under ... ------------------------------------------------------
        if cuda:
            self.side.wait_stream(main)
            with torch.cuda.stream(self.side):
                hv, segR, segS, seg_group, hR, hS = self._plan(cnt1, hist)
                hv_host = hv.to("cpu", non_blocking=True)
                side_done = torch.cuda.Event()
                side_done.record()
        else:
            hv, segR, segS, seg_group, hR, hS = self._plan(cnt1, hist)
            hv_host = hv
        # ---- 3. ... the routed pass-1 scatter into the send buffers (layout = local exclusive prefix) ------
        c = cnt1[:2 * F1].view(2, F1).to(torch.int64)
        loc = (torch.cumsum(c, 1) - c).to(torch.int32)
        be.shard_scatter(R, nR, loc[0], [sendR.data_ptr()] * G, 0)
        eR = self._event()
        be.shard_scatter(S, nS, loc[1], [sendS.data_ptr()] * G, 1)
        eS = self._event()
        if cuda:
            side_done.synchronize()            # the one host sync of the exchange
        host = [int(x) for x in hv_host.tolist()]
        if host[0] > self.capR or host[1] > self.capS:
            self.fallbacks += 1
            if cuda:
                main.wait_stream(self.side)
            out = ShardedJoin.run(self, R, S)
            out["exchange"] = "nccl-fallback"
            return out
        nR_recv, nS_recv = host[2], host[3]
        sendsR, sendsS = host[4:4 + G], host[4 + G:4 + 2 * G]
        dR, dS = host[4 + 2 * G:4 + 3 * G], host[4 + 3 * G:4 + 4 * G]
        # ---- 4. DMA exchange: 2G copy-engine transfers into the owners' buffers ------------------------------
        offs_r, offs_s, a, b = [], [], 0, 0
        for g in range(G):
            offs_r.append(a)
            offs_s.append(b)
            a += sendsR[g]
            b += sendsS[g]
        for k in range(G):
            g = (rank + k) % G                 # start with the local block, then walk the ring
            for which, (peers, send, offs, sends, dsts, ev) in enumerate(
                    ((self.peerR, sendR, offs_r, sendsR, dR, eR), (self.peerS, sendS, offs_s, sendsS, dS, eS))):
                self.exchange_copy(peers[g] + 8 * dsts[g], send.data_ptr() + 8 * offs[g], 8 * sends[g], 2 * k + which, ev)
        if cuda:
            for st in self.copy_streams:
                main.wait_stream(st)
            main.wait_stream(self.side)
        ex = self._event()
        flag = self._buf("flag", 2, torch.int32)
        dist.all_reduce(flag[:1], group=self.group)    # barrier: every rank's transfers have landed
        e2 = self._event()
        # ---- 5. local pass 2 + build/probe -------------------------------------------------------------------
        local = be.shard_join(_RawBuffer(self.peerR[rank]), nR_recv, segR, _RawBuffer(self.peerS[rank]),
                              nS_recv, segS, seg_group, G * per, per, self.b1, self.b2, hR, hS, self.bits)
        e3 = self._event()
        to_i64 = lambda v: v - (1 << 64) if v >= (1 << 63) else v
        res = torch.tensor([local["matches"], to_i64(local["checksum"]), to_i64(local["keysum"])], dtype=torch.int64,
                           device=self.device)
        dist.all_reduce(res, group=self.group)
        m, cs, ks = (int(x) for x in res.tolist())
        out = {"matches": m, "checksum": cs % (1 << 64), "keysum": ks % (1 << 64), "radix_bits": self.bits,
               "num_passes": 2, "bits_pass1": self.b1, "bits_pass2": self.b2, "tuples_sent": nR + nS,
               "tuples_kept": int(sendsR[rank] + sendsS[rank]), "ms_pass2": local.get("ms_pass2", 0.0),
               "ms_join": local.get("ms_join", 0.0), "exchange": "p2p-dma"}
        if e0 is not None:
            torch.cuda.synchronize()
            out["ms_hist"] = e0.elapsed_time(eh)
            out["ms_pass1"] = eh.elapsed_time(e2)      # scatter (sizing underneath) + transfers + barrier
            out["ms_sizing"] = 0.0                     # hidden under the scatter
            out["ms_scatter_kernels"] = eh.elapsed_time(eS)
            out["ms_exchange"] = eS.elapsed_time(e2)   # the part of the transfers not hidden under S's scatter
            out["ms_barrier"] = ex.elapsed_time(e2)
            out["ms_total"] = e0.elapsed_time(e3)
        return out

    def exchange_copy(self, dst_ptr, src_ptr, nbytes, lane, after):
        """one transfer of the exchange on copy stream `lane`, ordered after event `after`"""
        if not self.copy_streams:               # host stand-in backend (tests): synchronous copy
            self.backend.copy_async(dst_ptr, src_ptr, nbytes, None)
            return
        st = self.copy_streams[lane]
        st.wait_event(after)
        self.backend.copy_async(dst_ptr, src_ptr, nbytes, st.cuda_stream)


class MgShardedJoin:
    """The product multi-GPU path: the C host inside libb200aqp.so (csrc/mg.cu; NCCL + CUDA IPC from C++), the
    same code host/native_mg.cpp drives. Python only hands over the NCCL unique id (torch.distributed broadcast) and
    the device pointers. Region-layout exchange: no collective and no host read-back in front of the scatter."""

    def __init__(self, nR_total: int, nS_total: int, device, backend=None, group=None):
        import b200aqp as A
        self.A = A
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = device
        uid = torch.zeros(128, dtype=torch.uint8, device=device)
        if self.rank == 0:
            uid.copy_(torch.tensor(list(A.mg_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0, group=group)
        A.mg_init(self.rank, self.world, bytes(uid.cpu().tolist()), nR_total, nS_total)
        self.open = True

    def run(self, R: torch.Tensor, S: torch.Tensor) -> dict:
        torch.cuda.current_stream().synchronize()      # the C host runs on its own streams
        r = self.A.mg_join(R.data_ptr(), R.numel() // 2, S.data_ptr(), S.numel() // 2)
        r.update(num_passes=2, exchange="p2p-regions", ms_pass1=r["ms_scatter"] + r["ms_barrier"], ms_sizing=0.0,
                 ms_scatter_kernels=r["ms_scatter"], ms_exchange=0.0)
        return r

    def close(self):
        if self.open:
            self.A.mg_finalize()
            self.open = False
