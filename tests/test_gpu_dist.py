"""The real multi-GPU path: torchrun + NCCL, one process per GPU (needs >= 2 GPUs; skipped otherwise).
Launched by this test as a subprocess so it also works under plain `pytest -m gpu`."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu

SCRIPT = r'''
import os, sys
sys.path.insert(0, os.environ["AQP_ROOT"]); sys.path.insert(0, os.path.join(os.environ["AQP_ROOT"], "sgxv2-analytical-query-processing-benchmarks_b200"))
import torch, torch.distributed as dist
import b200aqp as A, b200aqp.dist as D
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr); A.init(lr)
dist.init_process_group("nccl", device_id=dev)
for logR, logS in ((20, 22), (24, 26)):
    nR, nS = 1 << logR, 1 << logS
    nRl, nSl = nR // world, nS // world
    R = torch.empty(2 * nRl, dtype=torch.int32, device=dev); S = torch.empty(2 * nSl, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    A.gen_pk_device(R.data_ptr(), nR, 11111, rank * nRl, nRl, st); A.gen_fk_device(S.data_ptr(), nS, nR, 22222, rank * nSl, nSl, st)
    rep = nS // nR
    for cls in (D.ShardedJoin, D.FusedShardedJoin, D.DmaShardedJoin):
        sj = cls(nR, nS, dev)
        for _ in range(3):
            o = sj.run(R, S)
        assert o["matches"] == nS, o
        assert o["keysum"] == rep * nR * (nR + 1) // 2, o
        assert o["checksum"] == rep * (nR * (nR - 1) // 2) + nS * (nS - 1) // 2, o
        if rank == 0: print("OK", cls.__name__, logR, logS, {k: round(v, 3) if isinstance(v, float) else v for k, v in o.items()})
        if hasattr(sj, "close"): sj.close()       # collective: peers unmap before anyone frees
        del sj
    # the C host (csrc/mg.cu): uniform and Zipf-skewed S (the hot key's owner receives far more than the mean:
    # worst-case regions, no overflow path), against the other variants' result on the same shards
    mgj = D.MgShardedJoin(nR, nS, dev)
    for _ in range(3):
        o = mgj.run(R, S)
    assert (o["matches"], o["keysum"]) == (nS, rep * nR * (nR + 1) // 2), o
    assert o["checksum"] == rep * (nR * (nR - 1) // 2) + nS * (nS - 1) // 2, o
    if rank == 0: print("OK MgShardedJoin", logR, logS, {k: round(v, 3) if isinstance(v, float) else v for k, v in o.items()})
    for z in (0.5, 1.0):
        A.gen_zipf_device(S.data_ptr(), nSl, nR, z, 22222, rank * nSl, st)
        torch.cuda.synchronize()
        ref = D.ShardedJoin(nR, nS, dev).run(R, S)
        o = mgj.run(R, S)
        assert o["matches"] == nS == ref["matches"] and o["checksum"] == ref["checksum"] and o["keysum"] == ref["keysum"], (z, o, ref)
        if rank == 0: print("OK MgShardedJoin zipf", z, logR, logS, round(o["ms_total"], 3))
    mgj.close()
# TPC-H Q12/Q3/Q19 sharded: every rank generates and filters its own rows, the joins are sharded joins (SURVEY 8e row 3)
for sf in (0.3, 2.0):
    A.tpch_generate_shard_device(sf, 5, rank, world)
    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0: uid.copy_(torch.tensor(list(A.mg_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, 0)
    A.tpch_mg_init(rank, world, bytes(uid.cpu().tolist()))
    got = {12: A.tpch_q12_mg(), 3: A.tpch_q3_mg(), 19: A.tpch_q19_mg()}
    A.mg_finalize()
    A.tpch_generate_device(sf, 5)                       # the whole data set on every rank: single-GPU answer
    for q in (12, 3, 19):
        want = A.tpch_query_device(q)
        assert got[q]["result_rows"] == want["result_rows"] and got[q]["join1_rows"] == want["join1_rows"], (sf, q, got[q], want)
        if rank == 0: print("OK tpch q%d mg" % q, sf, got[q]["result_rows"], round(got[q]["ms_total"], 3))
A.lib().b200_tpch_free_device()
dist.destroy_process_group()
'''


def test_nccl_sharded_join(tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else (4 if n < 8 else 8)
    script = tmp_path / "run_dist.py"
    script.write_text(SCRIPT)
    env = dict(os.environ, AQP_ROOT=ROOT)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29517", str(script)], env=env,
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert p.stdout.count("OK") == 18
