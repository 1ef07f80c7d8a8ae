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
    srcs = [os.path.join(HERE, f) for f in ("oracle_gen.c", "oracle_join.c", "oracle_scan.c", "oracle_tpch.c", "oracle.h")]
    stale = force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if os.path.isdir("/root/reference/Join-Benchmarks"):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])
        if os.path.exists(os.path.join(os.path.dirname(HERE), "sgxv2-analytical-query-processing-benchmarks_b200", "libb200aqp.so")):
            subprocess.check_call(["make", "-s", "-C", HERE, "dropin"])   # reference TPC-H pipelines linked against libb200aqp.so
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
        L.oracle_scan_sum.argtypes = [C.c_uint8, C.c_uint8, C.c_void_p, C.c_size_t]
        L.oracle_scan_sum.restype = C.c_uint64
        L.oracle_value_scan.argtypes = [C.c_uint8, C.c_uint8, C.c_void_p, C.c_size_t, C.c_void_p]
        L.oracle_value_scan.restype = C.c_uint64
        L.oracle_dict_code_range.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_dict_scan_8_64.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.oracle_dict_scan_8_64.restype = C.c_uint64
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


def scan_sum(lo, hi, data) -> int:
    return int(lib().oracle_scan_sum(lo, hi, _ptr(data), data.shape[0]))


def value_scan(lo, hi, data) -> np.ndarray:
    out = np.zeros(scan_count(lo, hi, data) + 64, dtype=np.uint32)
    n = lib().oracle_value_scan(lo, hi, _ptr(data), data.shape[0], _ptr(out))
    return out[:n]


def dict_code_range(lo: int, hi: int, dictionary: np.ndarray):
    a, b = np.zeros(1, dtype=np.uint8), np.zeros(1, dtype=np.uint8)
    d = np.ascontiguousarray(dictionary, dtype=np.int64)
    lib().oracle_dict_code_range(lo, hi, _ptr(d), _ptr(a), _ptr(b))
    return int(a[0]), int(b[0])


def dict_scan_8_64(lo: int, hi: int, dictionary: np.ndarray, data) -> np.ndarray:
    d = np.ascontiguousarray(dictionary, dtype=np.int64)
    assert d.shape[0] == 256
    out = np.zeros(data.shape[0] + 64, dtype=np.int64)
    n = lib().oracle_dict_scan_8_64(lo, hi, _ptr(d), _ptr(data), data.shape[0], _ptr(out))
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
        os.path.exists(os.path.join(REF_DIR, f))
        for f in ("libref_join.so", "libref_join_1p.so", "libref_scan.so", "libref_tpch.so"))


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
        L.ref_scan_sum.argtypes = [C.c_uint8, C.c_uint8, C.c_void_p, C.c_size_t]
        L.ref_scan_sum.restype = C.c_uint64
        L.ref_value_scan.argtypes = [C.c_uint8, C.c_uint8, C.c_void_p, C.c_size_t, C.c_void_p]
        L.ref_value_scan.restype = C.c_uint64
        L.ref_dict_scan_8_64.argtypes = [C.c_int, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
                                         C.c_uint64]
        L.ref_dict_scan_8_64.restype = C.c_uint64
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


def ref_scan_sum(lo, hi, data):
    assert data.ctypes.data % 64 == 0
    return int(ref_scan().ref_scan_sum(lo, hi, _ptr(data), data.shape[0]))


def ref_value_scan(lo, hi, data):
    assert data.ctypes.data % 64 == 0
    out = np.zeros(ref_scan_count(lo, hi, data) + 64, dtype=np.uint32)
    c = ref_scan().ref_value_scan(lo, hi, _ptr(data), data.shape[0], _ptr(out))
    return out[:c]


def ref_dict_scan_8_64(lo, hi, dictionary, data, which: int = 0):
    """which: 0 = dict_scan_8bit_64bit, 1 = ..._scalar_gather_scatter, 2 = ..._scalar_unroll, 3 = ..._opt_write"""
    assert data.ctypes.data % 64 == 0
    d = aligned_i64(256)
    d[:] = dictionary
    cap = data.shape[0] + 64
    out = np.zeros(cap, dtype=np.int64)
    c = ref_scan().ref_dict_scan_8_64(which, lo, hi, _ptr(d), _ptr(data), data.shape[0], _ptr(out), cap)
    return out[:c]


def aligned_i64(n: int, align: int = 64) -> np.ndarray:
    raw = np.empty(n * 8 + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + n * 8].view(np.int64)


def aligned(n: int, dtype, align: int = 64) -> np.ndarray:
    """n elements of dtype in 64-byte aligned memory (the reference's kernels use aligned 512-bit loads)"""
    item = np.dtype(dtype).itemsize
    raw = np.empty(n * item + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + n * item].view(dtype)


# ---- explicit-index scan and the 16- / 32-bit dictionary scans: restatement and compiled reference
def explicit_index_scan(lo, hi, index, data) -> np.ndarray:
    L = lib()
    L.oracle_explicit_index_scan.restype = C.c_uint64
    L.oracle_explicit_index_scan.argtypes = [C.c_uint8, C.c_uint8, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    out = np.zeros(data.shape[0] + 64, dtype=np.uint64)
    n = L.oracle_explicit_index_scan(lo, hi, _ptr(index), _ptr(data), data.shape[0], _ptr(out))
    return out[:n]


def wide_code_range(lo: int, hi: int, dictionary: np.ndarray):
    L = lib()
    L.oracle_wide_code_range.restype = None
    L.oracle_wide_code_range.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
    a, b = np.zeros(1, dtype=np.uint32), np.zeros(1, dtype=np.uint32)
    L.oracle_wide_code_range(lo, hi, _ptr(dictionary), dictionary.shape[0], _ptr(a), _ptr(b))
    return int(a[0]), int(b[0])


def dict_scan_wide(bits: int, lo: int, hi: int, dictionary: np.ndarray, data: np.ndarray) -> np.ndarray:
    L = lib()
    out = np.zeros(data.shape[0] + 64, dtype=np.int64)
    d = np.ascontiguousarray(dictionary, dtype=np.int64)
    if bits == 16:
        assert data.dtype == np.uint16 and d.shape[0] == 1 << 16
        L.oracle_dict_scan_16_64.restype = C.c_uint64
        L.oracle_dict_scan_16_64.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        n = L.oracle_dict_scan_16_64(lo, hi, _ptr(d), _ptr(data), data.shape[0], _ptr(out))
    else:
        assert data.dtype == np.uint32
        L.oracle_dict_scan_32_64.restype = C.c_uint64
        L.oracle_dict_scan_32_64.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]
        n = L.oracle_dict_scan_32_64(lo, hi, _ptr(d), d.shape[0], _ptr(data), data.shape[0], _ptr(out))
    return out[:n]


def ref_explicit_index_scan(lo, hi, index, data) -> np.ndarray:
    L = ref_scan()
    L.ref_explicit_index_scan.restype = C.c_uint64
    L.ref_explicit_index_scan.argtypes = [C.c_uint8, C.c_uint8, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    out = np.zeros(data.shape[0] + 64, dtype=np.uint64)
    n = L.ref_explicit_index_scan(lo, hi, _ptr(index), _ptr(data), data.shape[0], _ptr(out))
    return out[:n]


def ref_dict_scan_wide(bits: int, lo: int, hi: int, dictionary: np.ndarray, data: np.ndarray) -> np.ndarray:
    L = ref_scan()
    d = aligned(dictionary.shape[0], np.int64)
    d[:] = dictionary
    cap = data.shape[0] + 64
    out = np.zeros(cap, dtype=np.int64)
    if bits == 16:
        L.ref_dict_scan_16_64.restype = C.c_uint64
        L.ref_dict_scan_16_64.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_uint64]
        c = L.ref_dict_scan_16_64(lo, hi, _ptr(d), _ptr(data), data.shape[0], _ptr(out), cap)
    else:
        L.ref_dict_scan_32_64.restype = C.c_uint64
        L.ref_dict_scan_32_64.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p,
                                          C.c_uint64]
        c = L.ref_dict_scan_32_64(lo, hi, _ptr(d), d.shape[0], _ptr(data), data.shape[0], _ptr(out), cap)
    return out[:c]


# ----------------------------------------------------------------------------- TPC-H-style pipelines
class _LineItem(C.Structure):
    _fields_ = [("n", C.c_uint64), ("l_orderkey", C.c_void_p), ("l_shipdate", C.c_void_p), ("l_commitdate", C.c_void_p),
                ("l_receiptdate", C.c_void_p), ("l_shipmode", C.c_void_p), ("l_partkey", C.c_void_p),
                ("l_quantity", C.c_void_p), ("l_shipinstruct", C.c_void_p), ("l_returnflag", C.c_void_p)]


class _Orders(C.Structure):
    _fields_ = [("n", C.c_uint64), ("o_orderkey", C.c_void_p), ("o_orderdate", C.c_void_p), ("o_custkey", C.c_void_p)]


class _Customer(C.Structure):
    _fields_ = [("n", C.c_uint64), ("c_custkey", C.c_void_p), ("c_mktsegment", C.c_void_p), ("c_nationkey", C.c_void_p)]


class _Part(C.Structure):
    _fields_ = [("n", C.c_uint64), ("p_partkey", C.c_void_p), ("p_brand", C.c_void_p), ("p_size", C.c_void_p),
                ("p_container", C.c_void_p)]


_TPCH_DTYPES = {
    "lineitem": [("l_orderkey", ROW), ("l_shipdate", np.uint64), ("l_commitdate", np.uint64), ("l_receiptdate", np.uint64),
                 ("l_shipmode", np.uint8), ("l_partkey", np.uint32), ("l_quantity", np.float32),
                 ("l_shipinstruct", np.uint8), ("l_returnflag", np.int8)],
    "orders": [("o_orderkey", ROW), ("o_orderdate", np.uint64), ("o_custkey", np.uint32)],
    "customer": [("c_custkey", ROW), ("c_mktsegment", np.uint8), ("c_nationkey", np.uint32)],
    "part": [("p_partkey", ROW), ("p_brand", np.uint8), ("p_size", np.uint32), ("p_container", np.uint8)],
}
_TPCH_STRUCT = {"lineitem": _LineItem, "orders": _Orders, "customer": _Customer, "part": _Part}


def tpch_struct(name: str, cols: dict):
    """ctypes struct (layout of TpcHTypes.hpp) over a dict of numpy columns; 64-byte aligned copies are made
    because the reference's AVX-512 filters use aligned stream loads (Q12Predicates.hpp:63-70)."""
    keep = {}
    st = _TPCH_STRUCT[name]()
    n = None
    for col, dt in _TPCH_DTYPES[name]:
        a = cols[col]
        n = len(a) if n is None else n
        assert len(a) == n and a.dtype == np.dtype(dt), (col, a.dtype)
        raw = np.empty(a.nbytes + 128, dtype=np.uint8)
        off = (-raw.ctypes.data) % 64
        al = raw[off:off + a.nbytes].view(a.dtype)
        al[:] = a
        keep[col] = (raw, al)
        setattr(st, col, al.ctypes.data)
    st.n = n
    st._keep = keep
    return st


def synth_tpch(sf: float, seed: int = 1) -> dict:
    """numpy TPC-H-like tables with the encodings of the reference's loader (dictionary codes, epoch-second
    dates, sparse order keys). For CPU tests only; the GPU generator (b200_tpch_generate_device) has the same
    distributions but other random numbers."""
    rng = np.random.default_rng(seed)
    nc, no, np_ = max(3, int(150000 * sf)), max(1, int(1500000 * sf)), max(1, int(200000 * sf))
    nl = no * 4
    day = 86400
    t = {}
    ck = np.zeros(nc, dtype=ROW)
    ck["key"] = np.arange(1, nc + 1)
    ck["payload"] = np.arange(nc)
    t["customer"] = {"c_custkey": ck, "c_mktsegment": (rng.integers(0, 5, nc) == 0).astype(np.uint8),
                     "c_nationkey": rng.integers(0, 25, nc).astype(np.uint32)}
    ok = np.zeros(no, dtype=ROW)
    i = np.arange(no, dtype=np.uint64)
    ok["key"] = ((i >> 3) * 32 + (i & 7) + 1).astype(np.uint32)
    ok["payload"] = i.astype(np.uint32)
    od = (694224000 + day * rng.integers(0, 2406, no)).astype(np.uint64)
    k = rng.integers(0, nc - nc // 3, no)
    t["orders"] = {"o_orderkey": ok, "o_orderdate": od, "o_custkey": (k + k // 2 + 1).astype(np.uint32)}
    lk = np.zeros(nl, dtype=ROW)
    lk["key"] = np.repeat(ok["key"], 4)
    lk["payload"] = np.arange(nl)
    odl = np.repeat(od, 4)
    sd = odl + (day * rng.integers(1, 122, nl)).astype(np.uint64)
    mode = rng.integers(0, 7, nl)
    t["lineitem"] = {"l_orderkey": lk, "l_shipdate": sd,
                     "l_commitdate": odl + (day * rng.integers(30, 91, nl)).astype(np.uint64),
                     "l_receiptdate": sd + (day * rng.integers(1, 31, nl)).astype(np.uint64),
                     "l_shipmode": np.select([mode == 5, mode == 3, mode == 1], [1, 2, 3], 0).astype(np.uint8),
                     "l_partkey": rng.integers(1, np_ + 1, nl).astype(np.uint32),
                     "l_quantity": rng.integers(1, 51, nl).astype(np.float32),
                     "l_shipinstruct": (rng.integers(0, 4, nl) == 0).astype(np.uint8),
                     "l_returnflag": rng.choice(np.array([ord("R"), ord("A"), ord("N")], dtype=np.int8), nl)}
    pk = np.zeros(np_, dtype=ROW)
    pk["key"] = np.arange(1, np_ + 1)
    pk["payload"] = np.arange(np_)
    mn = rng.integers(1, 6, np_) * 10 + rng.integers(1, 6, np_)
    s1, s2 = rng.integers(0, 5, np_), rng.integers(0, 8, np_)
    cont = np.zeros(np_, dtype=np.uint8)
    for code, (a, b) in enumerate([(0, 0), (0, 1), (0, 5), (0, 4), (2, 2), (2, 1), (2, 4), (2, 5), (1, 0), (1, 1), (1, 5), (1, 4)], 1):
        cont[(s1 == a) & (s2 == b)] = code
    t["part"] = {"p_partkey": pk, "p_brand": np.select([mn == 12, mn == 23, mn == 34], [1, 2, 3], 0).astype(np.uint8),
                 "p_size": rng.integers(1, 51, np_).astype(np.uint32), "p_container": cont}
    return t


def _tpch_lib_setup(L):
    if getattr(L, "_tpch_ready", False):
        return
    L.oracle_tpch_q12.argtypes = [C.c_void_p, C.c_void_p, _u64p]
    L.oracle_tpch_q12.restype = C.c_int64
    L.oracle_tpch_q3.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, _u64p, _u64p]
    L.oracle_tpch_q3.restype = C.c_int64
    L.oracle_tpch_q19.argtypes = [C.c_void_p, C.c_void_p, _u64p, _u64p]
    L.oracle_tpch_q19.restype = C.c_int64
    L._tpch_ready = True


def tpch_query(q: int, tables: dict) -> dict:
    """Run the oracle restatement of Q3 / Q12 / Q19 on a dict of tables (dicts of numpy columns)."""
    L = lib()
    _tpch_lib_setup(L)
    st = {k: tpch_struct(k, v) for k, v in tables.items()}
    f = (C.c_uint64 * 3)()
    j1 = C.c_uint64(0)
    if q == 12:
        r = L.oracle_tpch_q12(C.byref(st["lineitem"]), C.byref(st["orders"]), f)
    elif q == 3:
        r = L.oracle_tpch_q3(C.byref(st["customer"]), C.byref(st["orders"]), C.byref(st["lineitem"]), f, C.byref(j1))
    elif q == 19:
        r = L.oracle_tpch_q19(C.byref(st["lineitem"]), C.byref(st["part"]), f, C.byref(j1))
    else:
        raise ValueError(q)
    return {"result_rows": int(r), "filtered": [int(x) for x in f], "join1_rows": int(j1.value)}


_ref_tpch = None


def ref_tpch():
    global _ref_tpch
    if _ref_tpch is None:
        L = C.CDLL(os.path.join(REF_DIR, "libref_tpch.so"))
        L.ref_tpch_q3.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, _f64p]
        L.ref_tpch_q3.restype = C.c_longlong
        L.ref_tpch_q12.argtypes = [C.c_void_p, C.c_void_p, C.c_int, _f64p]
        L.ref_tpch_q12.restype = C.c_longlong
        L.ref_tpch_q19.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_longlong), _f64p]
        L.ref_tpch_q19.restype = C.c_longlong
        _ref_tpch = L
    return _ref_tpch


def ref_tpch_query(q: int, tables: dict, nthreads: int = 4) -> dict:
    """The unmodified reference pipeline (oracle/_ref/libref_tpch.so) on the same tables."""
    L = ref_tpch()
    st = {k: tpch_struct(k, v) for k, v in tables.items()}
    sec = C.c_double(0)
    out = {}
    if q == 12:
        r = L.ref_tpch_q12(C.byref(st["lineitem"]), C.byref(st["orders"]), nthreads, C.byref(sec))
    elif q == 3:
        r = L.ref_tpch_q3(C.byref(st["customer"]), C.byref(st["orders"]), C.byref(st["lineitem"]), nthreads, C.byref(sec))
    elif q == 19:
        j = C.c_longlong(0)
        r = L.ref_tpch_q19(C.byref(st["lineitem"]), C.byref(st["part"]), nthreads, C.byref(j), C.byref(sec))
        out["join1_rows"] = int(j.value)
    else:
        raise ValueError(q)
    out.update(result_rows=int(r), seconds=sec.value)
    return out


# ---- link-level drop-in proof (tests/test_gpu_dropin.py) ----------------------------------------------------------
_dropin_tpch = None


def have_dropin() -> bool:
    return host_has_avx512() and os.path.exists(os.path.join(REF_DIR, "libdropin_tpch.so"))


def dropin_tpch_query(q: int, tables: dict, nthreads: int = 4) -> dict:
    """The reference's UNMODIFIED tpch_q3 / q12 / q19 (compiled from /root/reference into oracle/_ref/
    libdropin_tpch.so) running on libb200aqp.so's run_join / destroy_table instead of the reference's join library.
    Needs a GPU: the joins inside run on the device."""
    global _dropin_tpch
    if _dropin_tpch is None:
        L = C.CDLL(os.path.join(REF_DIR, "libdropin_tpch.so"))
        L.ref_tpch_q3.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, _f64p]
        L.ref_tpch_q3.restype = C.c_longlong
        L.ref_tpch_q12.argtypes = [C.c_void_p, C.c_void_p, C.c_int, _f64p]
        L.ref_tpch_q12.restype = C.c_longlong
        L.ref_tpch_q19.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_longlong), _f64p]
        L.ref_tpch_q19.restype = C.c_longlong
        _dropin_tpch = L
    L = _dropin_tpch
    st = {k: tpch_struct(k, v) for k, v in tables.items()}
    sec = C.c_double(0)
    out = {}
    if q == 12:
        r = L.ref_tpch_q12(C.byref(st["lineitem"]), C.byref(st["orders"]), nthreads, C.byref(sec))
    elif q == 3:
        r = L.ref_tpch_q3(C.byref(st["customer"]), C.byref(st["orders"]), C.byref(st["lineitem"]), nthreads, C.byref(sec))
    elif q == 19:
        j = C.c_longlong(0)
        r = L.ref_tpch_q19(C.byref(st["lineitem"]), C.byref(st["part"]), nthreads, C.byref(j), C.byref(sec))
        out["join1_rows"] = int(j.value)
    else:
        raise ValueError(q)
    out.update(result_rows=int(r), seconds=sec.value)
    return out
