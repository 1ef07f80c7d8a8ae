"""GPU parity of the sharded-join stage kernels (b200_shard_pass1_device / b200_shard_join_device).
The G ranks are emulated one after another on ONE GPU — the exchange is done with tensor slicing instead
of NCCL — so this runs in the single-GPU tier; tests/test_gpu_dist.py covers the real NCCL path."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _emulate(gpu, oracle, R, S, G, plan=None):
    import torch
    import b200aqp.dist as D
    dev = torch.device("cuda:0")
    be = D.CudaBackend()
    nR, nS = len(R), len(S)
    bits, b1, b2 = D.plan_bits(nR, G, (lambda n: plan) if plan else be.join_plan)
    lg = D.log2_exact(G)
    F1, P = 1 << b1, 1 << bits
    send, offs, hists = [], [], []
    for r in range(G):
        per_rank = []
        for rel, n in ((R, nR), (S, nS)):
            lo, hi = r * n // G, (r + 1) * n // G
            t = torch.from_numpy(rel[lo:hi].copy().view(np.int32)).to(dev)
            out = torch.empty(2 * (hi - lo) + 4, dtype=torch.int32, device=dev)
            hist = torch.zeros(P, dtype=torch.int32, device=dev)
            off = torch.zeros(F1 + 1, dtype=torch.int32, device=dev)
            be.shard_pass1(t, hi - lo, bits, b1, lg, out, hist, off)
            torch.cuda.synchronize()
            # pass-1 parity on this shard: partition p holds exactly the tuples with routed digit p
            o = out[:2 * (hi - lo)].cpu().numpy().view(oracle.ROW)
            offn = off.cpu().numpy().astype(np.int64)
            key = rel[lo:hi]["key"]
            p1 = key & (F1 - 1)
            routed = ((p1 >> lg) | (p1 << (b1 - lg))) & (F1 - 1) if lg else p1
            assert np.array_equal(np.diff(offn), np.bincount(routed, minlength=F1))
            got_p1 = o["key"] & (F1 - 1)
            got_routed = ((got_p1 >> lg) | (got_p1 << (b1 - lg))) & (F1 - 1) if lg else got_p1
            assert np.array_equal(got_routed, np.repeat(np.arange(F1), np.diff(offn)))
            assert np.array_equal(np.sort(o, order=["key", "payload"]), np.sort(rel[lo:hi], order=["key", "payload"]))
            per_rank.append((out, off, hist, hi - lo))
        send.append(per_rank)
    tot = {"matches": 0, "checksum": 0, "keysum": 0}
    for r in range(G):
        args = []
        for k in (0, 1):
            counts_all = torch.stack([(send[s][k][1][1:] - send[s][k][1][:-1]).to(torch.int64) for s in range(G)])
            hist_global = sum(send[s][k][2].to(torch.int64) for s in range(G)).to(torch.int32)
            _, recv, seg_off, seg_group = D.exchange_plan(counts_all, r, G)
            per = F1 // G
            parts = []
            for s in range(G):   # what the all-to-all would deliver from source s
                o = send[s][k][1].to(torch.int64)
                a, b = int(o[r * per]), int(o[(r + 1) * per])
                parts.append(send[s][k][0][2 * a:2 * b])
            buf = torch.cat(parts + [torch.zeros(4, dtype=torch.int32, device=dev)])
            assert int(recv.sum()) == (buf.numel() - 4) // 2
            args.append((buf, int(recv.sum()), seg_off.to(torch.int32), D.final_hist_slice(hist_global, r, G, b1, b2)))
        st = be.shard_join(args[0][0], args[0][1], args[0][2], args[1][0], args[1][1], args[1][2], seg_group, G * per,
                           per, b1, b2, args[0][3], args[1][3], bits)
        for k in tot:
            tot[k] += st[k]
        # the asynchronous form leaves the same three numbers on the device and reports the phase times later
        res = torch.full((4,), -1, dtype=torch.int64, device=dev)
        be.shard_join_async(args[0][0], args[0][1], args[0][2], args[1][0], args[1][1], args[1][2], seg_group, G * per,
                            per, b1, b2, args[0][3], args[1][3], bits, res)
        torch.cuda.synchronize()
        got = [int(x) % (1 << 64) for x in res[:3].tolist()]
        assert got == [st["matches"], st["checksum"], st["keysum"]] and int(res[3]) == -1
        times = be.shard_join_times()
        assert times["ms_total"] > 0 and times["ms_total"] >= times["ms_join"]
    exp = oracle.rho(R, S)
    assert (tot["matches"], tot["checksum"] % (1 << 64), tot["keysum"] % (1 << 64)) == \
        (exp["matches"], exp["checksum"], exp["keysum"])


@pytest.mark.parametrize("G", [1, 2, 4, 8])
def test_sharded_stages_uniform(gpu, oracle, G):
    R = oracle.set_rowid_payload(oracle.gen_pk(1 << 18, 11111))
    S = oracle.set_rowid_payload(oracle.gen_fk(1 << 20, 1 << 18, 22222))
    _emulate(gpu, oracle, R, S, G)


@pytest.mark.parametrize("G,plan", [(2, (6, 3, 3)), (4, (9, 4, 5)), (8, (3, 3, 0)), (2, (0, 0, 0)), (4, (14, 7, 7))])
def test_sharded_stages_plans_and_ragged_sizes(gpu, oracle, G, plan):
    R = oracle.set_rowid_payload(oracle.gen_pk(100003, 11111))
    S = oracle.set_rowid_payload(oracle.gen_fk(400009, 100003, 22222))
    _emulate(gpu, oracle, R, S, G, plan)


def test_sharded_stages_skew_and_misses(gpu, oracle):
    R = oracle.set_rowid_payload(oracle.gen_pk(1 << 16, 11111))
    S = oracle.set_rowid_payload(oracle.gen_zipf(1 << 19, 1 << 17, 1.0, seed=3))   # half of the key domain misses R
    _emulate(gpu, oracle, R, S, 4)


def _emulate_fused(gpu, oracle, R, S, G):
    """The fused scatter+exchange path with all G 'peer' receive buffers living on this one GPU."""
    import torch
    import b200aqp.dist as D
    dev = torch.device("cuda:0")
    be = D.CudaBackend()
    nR, nS = len(R), len(S)
    bits, b1, b2 = D.plan_bits(nR, G, be.join_plan)
    lg = D.log2_exact(G)
    F1, P = 1 << b1, 1 << bits
    per = F1 // G
    shards = [[torch.from_numpy(rel[r * n // G:(r + 1) * n // G].copy().view(np.int32)).to(dev)
               for rel, n in ((R, nR), (S, nS))] for r in range(G)]
    counts = [[None, None] for _ in range(G)]
    hists = [[None, None] for _ in range(G)]
    for r in range(G):
        for k in (0, 1):
            h = torch.zeros(P, dtype=torch.int32, device=dev)
            c = torch.zeros(F1, dtype=torch.int32, device=dev)
            be.shard_hist(shards[r][k], shards[r][k].numel() // 2, bits, b1, lg, h, c, k)
            torch.cuda.synchronize()
            counts[r][k], hists[r][k] = c.to(torch.int64), h
    recv = [[torch.zeros(2 * (n + 64), dtype=torch.int32, device=dev) for n in (nR, nS)] for _ in range(G)]
    for r in range(G):
        for k in (0, 1):
            counts_all = torch.stack([counts[s][k] for s in range(G)])
            h = torch.zeros(P, dtype=torch.int32, device=dev)
            c = torch.zeros(F1, dtype=torch.int32, device=dev)
            be.shard_hist(shards[r][k], shards[r][k].numel() // 2, bits, b1, lg, h, c, k)   # re-arm the slot for rank r
            be.shard_scatter(shards[r][k], shards[r][k].numel() // 2, D.dest_offsets(counts_all, r, G).to(torch.int32),
                             [recv[g][k].data_ptr() for g in range(G)], k)
    torch.cuda.synchronize()
    tot = {"matches": 0, "checksum": 0, "keysum": 0}
    for r in range(G):
        args = []
        for k in (0, 1):
            counts_all = torch.stack([counts[s][k] for s in range(G)])
            hist_global = sum(hists[s][k].to(torch.int64) for s in range(G)).to(torch.int32)
            _, rc, seg_off, seg_group = D.exchange_plan(counts_all, r, G)
            args.append((recv[r][k], int(rc.sum()), seg_off.to(torch.int32), D.final_hist_slice(hist_global, r, G, b1, b2)))
        st = be.shard_join(args[0][0], args[0][1], args[0][2], args[1][0], args[1][1], args[1][2], seg_group, G * per,
                           per, b1, b2, args[0][3], args[1][3], bits)
        for k in tot:
            tot[k] += st[k]
    exp = oracle.rho(R, S)
    assert (tot["matches"], tot["checksum"] % (1 << 64), tot["keysum"] % (1 << 64)) == \
        (exp["matches"], exp["checksum"], exp["keysum"])


@pytest.mark.parametrize("G", [1, 2, 4, 8])
def test_fused_scatter_exchange(gpu, oracle, G):
    R = oracle.set_rowid_payload(oracle.gen_pk(100003, 11111))
    S = oracle.set_rowid_payload(oracle.gen_fk(1 << 20, 100003, 22222))
    _emulate_fused(gpu, oracle, R, S, G)
    Sz = oracle.set_rowid_payload(oracle.gen_zipf(400009, 100003, 1.0, seed=5))
    _emulate_fused(gpu, oracle, R, Sz, G)


@pytest.mark.parametrize("G,b1,b2", [(2, 3, 3), (4, 7, 7), (8, 7, 7), (8, 3, 0), (1, 5, 6), (2, 8, 8)])
def test_exchange_plan_kernel_matches_tensor_formulation(gpu, G, b1, b2):
    """b200_exchange_plan_device (one launch) against the tensor formulation in b200aqp.dist that the CPU tests pin:
    received-segment tables, destination offsets, histogram slices and the sizes the host reads, for every rank."""
    import torch
    import b200aqp.dist as D
    dev = torch.device("cuda:0")
    be = D.CudaBackend()
    F1, P = 1 << b1, 1 << (b1 + b2)
    per = F1 // G
    gen = torch.Generator().manual_seed(1000 * G + b1)
    counts = torch.randint(0, 5000, (G, 2, F1), generator=gen, dtype=torch.int32)
    counts[0, 0, :per] = 0                                           # empty partitions too
    hist = torch.randint(0, 1 << 20, (2, P), generator=gen, dtype=torch.int32)
    c_dev, h_dev = counts.to(dev).contiguous(), hist.to(dev).contiguous()
    for rank in range(G):
        nseg = G * per
        seg = torch.full((2 * (nseg + 1),), -1, dtype=torch.int32, device=dev)
        dest = torch.full((2 * F1,), -1, dtype=torch.int32, device=dev)
        hsl = torch.full((2 * (per << b2),), -1, dtype=torch.int32, device=dev)
        hv = torch.zeros(6, dtype=torch.int64, device=dev)
        be.exchange_plan(c_dev, G, rank, b1, b2, h_dev, seg, dest, hsl, hv)
        torch.cuda.synchronize()
        c64 = counts.to(torch.int64)
        for rel in range(2):
            ca = c64[:, rel, :]
            _, recv, seg_off, _ = D.exchange_plan(ca, rank, G)
            assert seg[rel * (nseg + 1):(rel + 1) * (nseg + 1)].cpu().tolist() == seg_off.tolist()
            assert dest[rel * F1:(rel + 1) * F1].cpu().tolist() == D.dest_offsets(ca, rank, G).tolist()
            exp_h = D.final_hist_slice(hist[rel], rank, G, b1, b2)
            n2 = per << b2
            assert torch.equal(hsl[rel * n2:(rel + 1) * n2].cpu(), exp_h)
            assert int(hv[rel]) == int(ca.view(G, G, per).sum((0, 2)).max())
            assert int(hv[2 + rel]) == int(recv.sum())
            assert int(hv[4 + rel]) == int(ca[rank, rank * per:(rank + 1) * per].sum())


@pytest.mark.parametrize("G", [2, 8])
def test_fused_scatter_exchange_heavy_skew(gpu, oracle, G):
    """128 pass-1 partitions (64-slot bins) and an S side in which one key holds half of the tuples and a second
    one a tenth: the owning bins overflow in every tile, so the peer scatter's direct-store path, its unaligned
    heads after an overflow and the line-aligned ring flush all run side by side."""
    nR = 1 << 20
    R = oracle.set_rowid_payload(oracle.gen_pk(nR, 11111))
    rng = np.random.default_rng(99)
    nS = (1 << 21) + 12345
    S = oracle.gen_fk(nS, nR, 22222)[:nS].copy()
    u = rng.random(nS)
    S["key"] = np.where(u < 0.5, 777, np.where(u < 0.6, 4242, S["key"]))
    S = oracle.set_rowid_payload(S)
    _emulate_fused(gpu, oracle, R, S, G)
