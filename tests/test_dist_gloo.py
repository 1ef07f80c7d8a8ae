"""Multi-GPU host logic (b200aqp.dist) on CPU: world_size-2 and -4 gloo process groups with a numpy
stand-in for the two CUDA stage calls. Checks the routing order, split sizes, received-segment tables,
histogram slices and the final all-reduce against the oracle's single-process join."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import PKG, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class NumpyBackend:
    """Host stand-in for b200_shard_pass1_device / b200_shard_join_device (same contracts, see
    include/aqp/b200_aqp.h) so the torch.distributed logic can run without a GPU. Test double only."""

    def __init__(self, oracle, plan):
        self.O = oracle
        self.plan = plan
        self.checked = {}

    def join_plan(self, nR):
        return self.plan

    @staticmethod
    def _rot(p, b1, lg):
        m = (1 << b1) - 1
        return ((p >> lg) | (p << (b1 - lg))) & m if lg else p

    def shard_pass1(self, rel, n, bits, b1, lg, send, hist, part1_off):
        a = rel.numpy()[:2 * n].view(self.O.ROW)
        d = a["key"] & ((1 << bits) - 1)
        p1 = self._rot(d & ((1 << b1) - 1), b1, lg)
        routed = (d & ~np.uint32((1 << b1) - 1)) | p1
        hist.numpy()[:1 << bits] = np.bincount(routed, minlength=1 << bits).astype(np.int32)
        order = np.argsort(p1, kind="stable")
        send.numpy()[:2 * n] = a[order].view(np.int32)
        off = np.zeros((1 << b1) + 1, dtype=np.int64)
        off[1:] = np.cumsum(np.bincount(p1, minlength=1 << b1))
        part1_off.numpy()[:] = off.astype(np.int32)

    def shard_join(self, R, nR, segR, S, nS, segS, seg_group, nseg, ngroups, shift2, bits2, histR, histS, hash_shift):
        O = self.O
        r = R.numpy()[:2 * nR].view(O.ROW)
        s = S.numpy()[:2 * nS].view(O.ROW)
        lg, b1, rank = self.lg, shift2, self.rank
        for rel, seg, hist in ((r, segR.numpy(), histR.numpy()), (s, segS.numpy(), histS.numpy())):
            assert seg[0] == 0 and seg[-1] == len(rel) and len(seg) == nseg + 1
            got = np.zeros(ngroups << bits2, dtype=np.int64)
            for i in range(nseg):
                part = rel[seg[i]:seg[i + 1]]
                g = int(seg_group.numpy()[i])
                # every tuple of the segment is owned by this rank and lies in pass-1 partition group g
                assert ((part["key"] & ((1 << lg) - 1)) == rank).all()
                p1r = self._rot(part["key"] & ((1 << b1) - 1), b1, lg)
                assert (p1r == rank * ngroups + g).all()
                p2 = (part["key"] >> shift2) & ((1 << bits2) - 1)
                got[(g << bits2):((g + 1) << bits2)] += np.bincount(p2, minlength=1 << bits2)
            assert np.array_equal(got, hist.astype(np.int64))   # slice of the all-reduced histogram
        o = O.rho(np.ascontiguousarray(r), np.ascontiguousarray(s), nthreads=1)
        return {"matches": o["matches"], "checksum": o["checksum"], "keysum": o["keysum"]}


class _ShmBuffer:
    def __init__(self, shm):
        import ctypes as C
        self.shm = shm
        self.ptr = C.addressof(C.c_char.from_buffer(shm.buf))


class SharedNumpyBackend(NumpyBackend):
    """Adds host stand-ins for the peer-memory stage calls (b200_shard_hist_device, b200_shard_scatter_device,
    b200_copy_async, b200_ipc_export/open): receive buffers are POSIX shared memory mapped by every rank, the
    'device pointers' are this process's addresses of those mappings."""

    def __init__(self, oracle, plan):
        super().__init__(oracle, plan)
        self._slots = {}
        self._maps = []

    @staticmethod
    def _view(ptr, n):
        import ctypes as C
        return np.ctypeslib.as_array((C.c_int32 * (2 * n)).from_address(ptr))

    def alloc_shared(self, nbytes):
        from multiprocessing import shared_memory
        shm = shared_memory.SharedMemory(create=True, size=nbytes)
        name = shm.name.encode()
        assert len(name) < 64
        h = torch.zeros(64, dtype=torch.uint8)
        h[:len(name)] = torch.tensor(list(name), dtype=torch.uint8)
        self._maps.append(shm)
        return _ShmBuffer(shm), h

    def open_shared(self, handle):
        from multiprocessing import shared_memory
        name = bytes(handle.tolist()).rstrip(b"\0").decode()
        shm = shared_memory.SharedMemory(name=name)
        self._maps.append(shm)
        return _ShmBuffer(shm).ptr

    def close(self, unlink_own):
        for i, m in enumerate(self._maps):
            m.close()
        for m in unlink_own:
            m.unlink()

    def shard_hist(self, rel, n, bits, b1, lg, hist, counts1, slot):
        a = rel.numpy()[:2 * n].view(self.O.ROW)
        d = a["key"] & ((1 << bits) - 1)
        p1 = self._rot(d & ((1 << b1) - 1), b1, lg)
        routed = (d & ~np.uint32((1 << b1) - 1)) | p1
        hist.numpy()[:1 << bits] = np.bincount(routed, minlength=1 << bits).astype(np.int32)
        counts1.numpy()[:1 << b1] = np.bincount(p1, minlength=1 << b1).astype(np.int32)
        self._slots[slot] = (p1, b1, lg)

    def shard_scatter(self, rel, n, dest_off, dest_ptrs, slot):
        p1, b1, lg = self._slots[slot]
        a = rel.numpy()[:2 * n].view(np.int64)
        per_shift = b1 - lg
        off = dest_off.numpy()
        for p in np.unique(p1):
            rows = a[p1 == p]
            dst = self._view(dest_ptrs[int(p) >> per_shift] + 8 * int(off[p]), len(rows)).view(np.int64)
            dst[:] = rows

    def copy_async(self, dst_ptr, src_ptr, nbytes, stream):
        import ctypes as C
        C.memmove(dst_ptr, src_ptr, nbytes)

    def shard_join(self, R, nR, segR, S, nS, segS, *rest):
        wrap = lambda b, n: torch.from_numpy(self._view(b.data_ptr(), n).copy()) if not isinstance(b, torch.Tensor) else b
        return super().shard_join(wrap(R, nR), nR, segR, wrap(S, nS), nS, segS, *rest)


def _peer_worker(rank, world, port, nR, nS, plan, cls_name, ret):
    for p in (ROOT, PKG):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    import b200aqp.dist as D
    R = O.set_rowid_payload(O.gen_pk(nR, 11111))
    S = O.set_rowid_payload(O.gen_fk(nS, nR, 22222))
    exp = O.rho(R, S)
    be = SharedNumpyBackend(O, plan)
    be.rank, be.lg = rank, D.log2_exact(world)
    lo_r, hi_r = rank * nR // world, (rank + 1) * nR // world
    lo_s, hi_s = rank * nS // world, (rank + 1) * nS // world
    Rl = torch.from_numpy(R[lo_r:hi_r].copy().view(np.int32))
    Sl = torch.from_numpy(S[lo_s:hi_s].copy().view(np.int32))
    sj = getattr(D, cls_name)(nR, nS, torch.device("cpu"), backend=be)
    own = [b.shm for b in sj._own]
    try:
        for _ in range(2):                                          # buffers are reused across runs
            out = sj.run(Rl, Sl)
            assert (out["matches"], out["checksum"], out["keysum"]) == (exp["matches"], exp["checksum"], exp["keysum"]), out
            assert out["exchange"] == ("p2p-dma" if cls_name == "DmaShardedJoin" else "p2p-fused")
            dist.barrier()                                          # nobody overwrites a buffer a peer still reads
        # a capacity too small for the data must take the NCCL path on every rank, with the same answer
        sj.capR = sj.capS = 1
        out = sj.run(Rl, Sl)
        assert out["exchange"] == "nccl-fallback" and out["matches"] == exp["matches"] and sj.fallbacks == 1
        ret[rank] = out["matches"]
        dist.barrier()
    finally:
        del sj
        be.close(own)
    dist.destroy_process_group()


@pytest.mark.parametrize("cls_name", ["FusedShardedJoin", "DmaShardedJoin"])
@pytest.mark.parametrize("world,nR,nS,plan", [(2, 20011, 100003, (6, 3, 3)), (4, 30000, 120001, (7, 3, 4))])
def test_peer_memory_exchange_host_logic(cls_name, world, nR, nS, plan):
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_peer_worker, args=(world, port, nR, nS, plan, cls_name, ret), nprocs=world, join=True)
    assert len(ret) == world and len(set(ret.values())) == 1


def _worker(rank, world, port, nR, nS, plan, ret):
    for p in (ROOT, PKG):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    import b200aqp.dist as D
    R = O.set_rowid_payload(O.gen_pk(nR, 11111))
    S = O.set_rowid_payload(O.gen_fk(nS, nR, 22222))
    exp = O.rho(R, S)
    be = NumpyBackend(O, plan)
    be.rank, be.lg = rank, D.log2_exact(world)
    lo_r, hi_r = rank * nR // world, (rank + 1) * nR // world     # row-range shards (uneven on purpose)
    lo_s, hi_s = rank * nS // world, (rank + 1) * nS // world
    Rl = torch.from_numpy(R[lo_r:hi_r].copy().view(np.int32))
    Sl = torch.from_numpy(S[lo_s:hi_s].copy().view(np.int32))
    sj = D.ShardedJoin(nR, nS, torch.device("cpu"), backend=be)
    out = sj.run(Rl, Sl)
    assert (out["matches"], out["checksum"], out["keysum"]) == (exp["matches"], exp["checksum"], exp["keysum"]), out
    tot = torch.tensor([out["tuples_sent"]], dtype=torch.int64)
    dist.all_reduce(tot)
    assert int(tot.item()) == nR + nS
    out2 = sj.run(Rl, Sl)                                           # buffers are reused across runs
    assert out2["matches"] == exp["matches"]
    ret[rank] = out["matches"]
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nR,nS,plan", [(2, 20011, 100003, (6, 3, 3)), (2, 1 << 14, 1 << 16, (4, 4, 0)),
                                               (4, 30000, 120001, (7, 3, 4)), (2, 5000, 9999, (0, 0, 0))])
def test_sharded_join_host_logic(world, nR, nS, plan):
    port = _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, nR, nS, plan, ret), nprocs=world, join=True)
    assert len(ret) == world and len(set(ret.values())) == 1


def test_exchange_plan_and_hist_slice():
    sys.path.insert(0, PKG)
    import b200aqp.dist as D
    counts = torch.tensor([[1, 2, 3, 4, 5, 6, 7, 8], [10, 20, 30, 40, 50, 60, 70, 80]])   # 2 ranks, F1 = 8
    send, recv, seg_off, seg_group = D.exchange_plan(counts, rank=1, world=2)
    assert send.tolist() == [100, 260]                 # rank 1 sends partitions 0-3 to rank 0, 4-7 to itself
    assert recv.tolist() == [26, 260]                  # and receives partitions 4-7 from both
    assert seg_off.tolist() == [0, 5, 11, 18, 26, 76, 136, 206, 286]
    assert seg_group.tolist() == [0, 1, 2, 3, 0, 1, 2, 3]
    # fused exchange: where each rank's segments start inside the owners' buffers = the owners' seg_off
    assert D.dest_offsets(counts, 0, 2).tolist() == [0, 1, 3, 6, 0, 5, 11, 18]
    assert D.dest_offsets(counts, 1, 2).tolist() == [10, 20, 40, 70, 26, 76, 136, 206]
    for owner in (0, 1):
        seg = D.exchange_plan(counts, owner, 2)[2].tolist()
        for src in (0, 1):
            assert D.dest_offsets(counts, src, 2).view(2, 4)[owner].tolist() == seg[src * 4:src * 4 + 4]
    h = torch.arange(16)                               # b1 = 2, b2 = 2: index = p1 | p2 << 2
    assert D.final_hist_slice(h, rank=1, world=2, b1=2, b2=2).tolist() == [2, 6, 10, 14, 3, 7, 11, 15]
    assert D.plan_bits(1 << 27, 8, lambda n: (14, 7, 7)) == (14, 7, 7)
    assert D.plan_bits(5000, 4, lambda n: (0, 0, 0)) == (2, 2, 0)
    with pytest.raises(ValueError):
        D.log2_exact(6)
