"""C-ABI checks that need no GPU: the library loads, exports every symbol the header declares,
the structs have the reference's layout, the host-side generators reproduce the reference, and a
compute call without a GPU fails loudly instead of falling back."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import PKG, ROOT
from helpers import sha

HEADERS = [os.path.join(ROOT, "include", "aqp", h) for h in ("b200_aqp.h", "b200_tpch.h")]


def _declared_functions():
    src = "\n".join(open(h).read() for h in HEADERS)
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", src)
    return sorted(set(n for n in names if n not in ("defined",)))


def test_library_exports_every_declared_symbol(aqp):
    declared = _declared_functions()
    assert len(declared) >= 40
    L = C.CDLL(aqp.LIB_PATH)
    for name in declared:
        assert hasattr(L, name), f"{name} declared in b200_aqp.h but not exported"
        assert name in aqp.SYMBOLS, f"{name} has no ctypes signature in b200aqp"
    assert sorted(aqp.SYMBOLS) == declared
    nm = subprocess.run(["nm", "-D", "--defined-only", aqp.LIB_PATH], capture_output=True, text=True).stdout
    for name in declared:
        assert re.search(rf"\bT {name}\b", nm), f"{name} is not a defined text symbol"


def test_struct_layout_matches_reference(aqp, tmp_path):
    # ctypes mirrors == C header (compiled here) == reference sizes probed from data-types.h
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "aqp/data_types.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",sizeof(struct row_t),'
                   'sizeof(struct table_t),sizeof(struct output_triple_t),sizeof(struct table_chunk_t),'
                   'sizeof(struct chunked_table_t),sizeof(struct result_t),sizeof(struct joinconfig_t),'
                   '(size_t)TUPLES_PER_CHUNK,offsetof(struct result_t,result),offsetof(struct joinconfig_t,MATERIALIZE));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert got == [8, 24, 12, 16376, 40, 48, 48, 1364, 32, 32]
    assert [C.sizeof(aqp.Row), C.sizeof(aqp.Table), C.sizeof(aqp.ChunkedTable), C.sizeof(aqp.Result),
            C.sizeof(aqp.JoinConfig)] == [8, 24, 40, 48, 48]
    assert aqp.Result.result.offset == 32 and aqp.JoinConfig.MATERIALIZE.offset == 32
    assert aqp.TUPLES_PER_CHUNK == 1364 and aqp.CHUNK_BYTES == 16376
    if os.path.isdir("/root/reference"):   # container only: the reference header itself
        ref = tmp_path / "ref.cpp"
        ref.write_text('#include <cstdio>\n#include "data-types.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu\\n",'
                       'sizeof(row_t),sizeof(table_t),sizeof(output_triple_t),sizeof(table_chunk_t),'
                       'sizeof(chunked_table_t),sizeof(result_t),sizeof(joinconfig_t));}\n')
        subprocess.check_call(["g++", "-I/root/reference/Join-Benchmarks/lib/SharedHeaders/include", str(ref), "-o",
                               str(tmp_path / "ref")])
        assert [int(x) for x in subprocess.check_output([str(tmp_path / "ref")]).split()] == got[:7]


def test_host_generators_reproduce_reference(aqp, oracle, golden):
    for c in golden["generator"]:
        if c["n"] > (1 << 22):
            continue   # the big ones are covered by the oracle test; keep the CPU suite short
        rel = {"pk": lambda: aqp.host_gen_pk(c["n"], c["seed"]),
               "fk": lambda: aqp.host_gen_fk(c["n"], c["maxid"], c["seed"]),
               "fk_sel": lambda: aqp.host_gen_fk_sel(c["n"], c["maxid"], c["seed"])}[c["kind"]]()
        assert [int(x) for x in rel["key"][:8]] == c["first8"], c
        assert sha(rel["key"]) == c["sha256_keys"], c
    assert np.array_equal(aqp.host_gen_pk(12345, 77), oracle.gen_pk(12345, 77))


def test_host_zipf_is_seeded_and_skewed(aqp):
    a = aqp.host_gen_zipf(1 << 15, 1 << 10, 1.0, 5)
    b = aqp.host_gen_zipf(1 << 15, 1 << 10, 1.0, 5)
    assert np.array_equal(a, b)                                  # repeatable, unlike genzipf.cpp:44-45
    assert a["key"].min() >= 1 and a["key"].max() <= 1 << 10     # alphabet is 1..maxid
    top = np.bincount(a["key"]).max() / len(a)
    assert 0.10 < top < 0.17                                     # z=1, 1024 symbols: P(top) = 1/H_1024 = 0.133


def test_join_plan(aqp):
    assert aqp.join_plan(1 << 27) == (14, 7, 7)
    assert aqp.join_plan(1 << 24) == (11, 5, 6)
    assert aqp.join_plan(1 << 20) == (7, 7, 0)
    assert aqp.join_plan(8192) == (0, 0, 0)
    assert aqp.join_plan(8193) == (1, 1, 0)


def test_no_cpu_fallback_without_gpu(aqp):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = os.path.exists("/dev/nvidia0")
    if has_gpu:
        pytest.skip("GPU present")
    assert aqp.lib().b200_init(-1) != 0
    assert b"no CPU fallback" in aqp.lib().b200_last_error()
    with pytest.raises(aqp.AqpError):
        aqp.join_device(0, 0, 0, 0)


def test_unknown_algorithm_exits_like_reference(aqp):
    # joins.cpp:70-73: unknown algorithm -> error + exit(EXIT_FAILURE); run in a child process
    code = ("import sys; sys.path.insert(0, %r); import numpy as np, b200aqp as A; "
            "R=np.zeros(4,dtype=A.ROW); A.run_join(R,R,algorithm=b'PHT')" % PKG)
    p = subprocess.run(["python", "-c", code], capture_output=True, text=True)
    assert p.returncode == 1
    assert "Algorithm not found: PHT" in p.stderr
