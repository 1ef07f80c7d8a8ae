"""ctypes access to the CPU oracle (liboracle.so) and to the compiled reference (oracle/_ref/).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / ``--impl reference`` legs — never by the product package.  See oracle/oracle.h
for what each function restates (reference file:line) and how parity is pinned.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

ROW = np.dtype([("key", np.uint32), ("payload", np.uint32)])
TRIPLE = np.dtype([("key", np.uint32), ("Rpayload", np.uint32), ("Spayload", np.uint32)])

_u8p = C.POINTER(C.c_uint8)
_u64p = C.POINTER(C.c_uint64)
_i64p = C.POINTER(C.c_int64)
_f64p = C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    """Compile the C restatement (and, when /root/reference is present, oracle/_ref)."""
    so = os.path.join(HERE, "liboracle.so")
    srcs = [os.path.join(HERE, f) for f in ("oracle_gen.c", "oracle_join.c", "oracle_scan.c", "oracle.h")]
    stale = force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if os.path.isdir("/root/reference/Join-Benchmarks"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])
    return so


def _ptr(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.oracle_srand.argtypes = [C.c_uint32]
        L.oracle_rand.restype = C.c_int32
        L.oracle_gen_pk.argtypes = [C.c_void_p, C.c_uint64]
        L.oracle_gen_fk.argtypes = [C.c_void_p, C.c_uint64, C.c_int64]
        L.oracle_gen_fk_sel.argtypes = [C.c_void_p, C.c_uint64, C.c_int64]
        L.oracle_zipf_lut.argtypes = [C.c_void_p, C.c_uint32, C.c_double]
        L.oracle_zipf_pos.argtypes = [C.c_void_p, C.c_uint32, C.c_double]
        L.oracle_zipf_pos.restype = C.c_uint32
        L.oracle_calc_num_radix_bits.argtypes = [C.c_uint64, C.c_uint64]
        L.oracle_calc_num_radix_bits.restype = C.c_uint32
        L.oracle_calc_num_passes.argtypes = [C.c_uint32]
        L.oracle_calc_num_passes.restype = C.c_uint32
        L.oracle_radix_partition.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        L.oracle_bucket_chaining_join.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint32,
                                                  _u64p, _u64p, C.c_void_p, C.c_uint64, _u64p]
        L.oracle_bucket_chaining_join.restype = C.c_int64
        L.oracle_rho.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int, C.c_int,
                                 _u64p, _u64p, C.c_void_p, C.c_uint64]
        L.oracle_rho.restype = C.c_int64
        L.oracle_scan_count.argtypes = [C.c_uint8, C.c_uint8, C.c_void_p, C.c_size_t]
        L.oracle_scan_count.restype = C.c_uint64
        L.oracle_bitvector_scan.argtypes = [C.c_uint8, C.c_uint8, C.c_void_p, C.c_size_t, C.c_void_p]
        for f in (L.oracle_index_scan, L.oracle_scalar_index_scan):
            f.argtypes = [C.c_uint8, C.c_uint8, C.c_void_p, C.c_size_t, C.c_void_p]
            f.restype = C.c_uint64
        L.oracle_fill_tiled_column.argtypes = [C.c_void_p, C.c_size_t]
        _lib = L
    return _lib


# ----------------------------------------------------------------------------- generators
def gen_pk(n: int, seed: int) -> np.ndarray:
    rel = np.zeros(n, dtype=ROW)
    lib().oracle_srand(seed)
    lib().oracle_gen_pk(_ptr(rel), n)
    return rel


def gen_fk(n: int, maxid: int, seed: int) -> np.ndarray:
    rel = np.zeros(n, dtype=ROW)
    lib().oracle_srand(seed)
    lib().oracle_gen_fk(_ptr(rel), n, maxid)
    return rel


def gen_fk_sel(n: int, maxid: int, seed: int) -> np.ndarray:
    rel = np.zeros(n, dtype=ROW)
    lib().oracle_srand(seed)
    lib().oracle_gen_fk_sel(_ptr(rel), n, maxid)
    return rel


def gen_zipf(n: int, maxid: int, z: float, seed: int) -> np.ndarray:
    """Zipf-skewed FK relation with the reference's LUT/binary-search algorithm
    (genzipf.cpp:58-137) but a *seeded* numpy PRNG: the reference seeds from
    std::random_device (genzipf.cpp:44-45,:104-105) and is not reproducible,
    so parity on skewed inputs is always on shared bytes."""
    rng = np.random.default_rng(seed)
    alphabet = rng.permutation(maxid).astype(np.uint32) + 1
    lut = np.empty(maxid, dtype=np.float64)
    lib().oracle_zipf_lut(_ptr(lut), maxid, z)
    r = rng.random(n)
    # vectorised form of oracle_zipf_pos (first index with lut[idx] >= r, clamped)
    pos = np.minimum(np.searchsorted(lut, r, side="left"), maxid - 1)
    rel = np.zeros(n, dtype=ROW)
    rel["key"] = alphabet[pos]
    return rel


def set_rowid_payload(rel: np.ndarray) -> np.ndarray:
    """payload = row index (TPC-H loader convention, JB/App/TpcH/TpcHCommons.cpp:332,:413)."""
    rel["payload"] = np.arange(rel.shape[0], dtype=np.uint32)
    return rel


# ----------------------------------------------------------------------------- join
def rho(R: np.ndarray, S: np.ndarray, nthreads: int = 1, force_2_passes: bool = True, materialize: bool = False):
    cs, ks = C.c_uint64(0), C.c_uint64(0)
    out = None
    cap = 0
    if materialize:   # size the output with a count-only run first (duplicate build keys can exceed |S|)
        cap = int(lib().oracle_rho(_ptr(R), R.shape[0], _ptr(S), S.shape[0], nthreads, int(force_2_passes),
                                   None, None, None, 0)) + 16
        out = np.zeros(cap, dtype=TRIPLE)
    m = lib().oracle_rho(_ptr(R), R.shape[0], _ptr(S), S.shape[0], nthreads, int(force_2_passes),
                         C.byref(cs), C.byref(ks), _ptr(out) if out is not None else None, cap)
    res = {"matches": int(m), "checksum": int(cs.value), "keysum": int(ks.value)}
    if materialize:
        assert m <= cap
        res["triples"] = out[:m]
    return res


def radix_partition(rel: np.ndarray, shift: int, bits: int):
    out = np.zeros_like(rel)
    offs = np.zeros((1 << bits) + 1, dtype=np.uint64)
    lib().oracle_radix_partition(_ptr(rel), rel.shape[0], shift, bits, _ptr(out), _ptr(offs))
    return out, offs


# ----------------------------------------------------------------------------- scan
def tiled_column(n: int) -> np.ndarray:
    a = np.empty(n, dtype=np.uint8)
    lib().oracle_fill_tiled_column(_ptr(a), n)
    return a


def scan_count(lo, hi, data) -> int:
    return int(lib().oracle_scan_count(lo, hi, _ptr(data), data.shape[0]))


def bitvector_scan(lo, hi, data) -> np.ndarray:
    out = np.zeros(data.shape[0] // 64, dtype=np.uint64)
    lib().oracle_bitvector_scan(lo, hi, _ptr(data), data.shape[0], _ptr(out))
    return out


def index_scan(lo, hi, data) -> np.ndarray:
    out = np.zeros(scan_count(lo, hi, data) + 64, dtype=np.uint64)
    n = lib().oracle_index_scan(lo, hi, _ptr(data), data.shape[0], _ptr(out))
    return out[:n]


def scalar_index_scan(lo, hi, data) -> np.ndarray:
    out = np.zeros(data.shape[0] + 64, dtype=np.uint64)
    n = lib().oracle_scalar_index_scan(lo, hi, _ptr(data), data.shape[0], _ptr(out))
    return out[:n]


# ----------------------------------------------------------------------------- compiled reference
def host_has_avx512() -> bool:
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        return False
    return all(f in flags for f in ("avx512f", "avx512bw", "avx512vl", "avx512dq", "avx512cd", "avx512vbmi", "avx512_vbmi2"))


def have_ref() -> bool:
    return host_has_avx512() and all(
        os.path.exists(os.path.join(REF_DIR, f)) for f in ("libref_join.so", "libref_join_1p.so", "libref_scan.so"))


_ref_join = {}
_ref_scan = None


def ref_join(force_2_passes: bool = True):
    """The unmodified reference RHO + generators (oracle/_ref/libref_join[_1p].so)."""
    key = bool(force_2_passes)
    if key not in _ref_join:
        L = C.CDLL(os.path.join(REF_DIR, "libref_join.so" if key else "libref_join_1p.so"))
        L.ref_tsc_hz.restype = C.c_double
        L.ref_seed_generator.argtypes = [C.c_uint]
        L.ref_create_relation_pk.argtypes = [C.c_uint64]
        L.ref_create_relation_pk.restype = C.c_void_p
        for f in (L.ref_create_relation_fk, L.ref_create_relation_fk_sel):
            f.argtypes = [C.c_uint64, C.c_int64]
            f.restype = C.c_void_p
        L.ref_create_relation_zipf.argtypes = [C.c_uint64, C.c_int64, C.c_double]
        L.ref_create_relation_zipf.restype = C.c_void_p
        L.ref_free.argtypes = [C.c_void_p]
        L.ref_rho.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_int, C.c_int, _i64p, _u64p, _u64p,
                              C.c_void_p, C.c_uint64, _u64p, _f64p]
        L.ref_rho.restype = C.c_int
        _ref_join[key] = L
    return _ref_join[key]


def _take(L, p, n) -> np.ndarray:
    a = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint32)), shape=(n, 2)).copy().view(ROW).reshape(n)
    L.ref_free(p)
    return a


def ref_gen_pk(n, seed):
    L = ref_join()
    L.ref_seed_generator(seed)
    return _take(L, L.ref_create_relation_pk(n), n)


def ref_gen_fk(n, maxid, seed):
    L = ref_join()
    L.ref_seed_generator(seed)
    return _take(L, L.ref_create_relation_fk(n, maxid), n)


def ref_gen_fk_sel(n, maxid, seed):
    L = ref_join()
    L.ref_seed_generator(seed)
    return _take(L, L.ref_create_relation_fk_sel(n, maxid), n)


def ref_gen_zipf(n, maxid, z):
    L = ref_join()
    return _take(L, L.ref_create_relation_zipf(n, maxid, z), n)


PHASES = ("total", "partition", "pass1", "pass2", "buildprobe", "build", "probe")


def ref_rho(R, S, nthreads=1, materialize=False, force_2_passes=True):
    """Run the reference RHO(). Returns matches (+ checksum/keysum/triples when materialised),
    the reference's own phase cycle counters and seconds = Total Join Time / measured TSC Hz."""
    L = ref_join(force_2_passes)
    m, cs, ks, wall = C.c_int64(0), C.c_uint64(0), C.c_uint64(0), C.c_double(0)
    cyc = (C.c_uint64 * 7)()
    out, cap = None, 0
    if materialize:
        cap = int(S.shape[0]) * 2 + 16
        out = np.zeros(cap, dtype=TRIPLE)
    rc = L.ref_rho(_ptr(R), R.shape[0], _ptr(S), S.shape[0], nthreads, int(materialize), C.byref(m), C.byref(cs),
                   C.byref(ks), _ptr(out) if out is not None else None, cap, cyc, C.byref(wall))
    assert rc == 0
    hz = L.ref_tsc_hz()
    res = {"matches": int(m.value), "cycles": dict(zip(PHASES, [int(c) for c in cyc])), "tsc_hz": hz,
           "seconds": cyc[0] / hz, "wall_seconds": wall.value}
    if materialize:
        res.update(checksum=int(cs.value), keysum=int(ks.value), triples=out[:m.value])
    return res


def ref_scan():
    global _ref_scan
    if _ref_scan is None:
        L = C.CDLL(os.path.join(REF_DIR, "libref_scan.so"))
        L.ref_scan_count.argtypes = [C.c_uint8, C.c_uint8, C.c_void_p, C.c_size_t]
        L.ref_scan_count.restype = C.c_uint64
        L.ref_bitvector_scan.argtypes = [C.c_uint8, C.c_uint8, C.c_void_p, C.c_size_t, C.c_void_p]
        L.ref_index_scan.argtypes = [C.c_uint8, C.c_uint8, C.c_void_p, C.c_size_t, C.c_void_p]
        L.ref_index_scan_self_alloc.argtypes = [C.c_uint8, C.c_uint8, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint64]
        L.ref_index_scan_self_alloc.restype = C.c_uint64
        L.ref_scan_mt.argtypes = [C.c_int, C.c_uint8, C.c_uint8, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                  C.c_void_p]
        L.ref_scan_mt.restype = C.c_double
        _ref_scan = L
    return _ref_scan


def aligned_u8(n: int, align: int = 64) -> np.ndarray:
    """The reference scan kernels use aligned 512-bit loads; its columns are 64-B aligned
    (SimdScanMulti/App/App.cpp:72)."""
    raw = np.empty(n + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + n]


def ref_bitvector_scan(lo, hi, data):
    assert data.ctypes.data % 64 == 0
    out = np.zeros(data.shape[0] // 64, dtype=np.uint64)
    ref_scan().ref_bitvector_scan(lo, hi, _ptr(data), data.shape[0], _ptr(out))
    return out


def ref_scan_count(lo, hi, data):
    assert data.ctypes.data % 64 == 0
    return int(ref_scan().ref_scan_count(lo, hi, _ptr(data), data.shape[0]))


def ref_index_scan(lo, hi, data):
    assert data.ctypes.data % 64 == 0
    cnt = ref_scan_count(lo, hi, data)
    out = np.zeros(cnt + 64, dtype=np.uint64)
    ref_scan().ref_index_scan(lo, hi, _ptr(data), data.shape[0], _ptr(out))
    return out[:cnt]


def ref_index_scan_self_alloc(lo, hi, data):
    assert data.ctypes.data % 64 == 0
    cap = data.shape[0] + 64
    out = np.zeros(cap, dtype=np.uint64)
    c = ref_scan().ref_index_scan_self_alloc(lo, hi, _ptr(data), data.shape[0], _ptr(out), cap)
    return out[:c]
