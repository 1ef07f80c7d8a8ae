"""b200aqp — thin ctypes binding of libb200aqp.so (include/aqp/b200_aqp.h).

This is plumbing for tests/ and bench.py, not the product: every function here forwards to one
extern "C" entry point of the CUDA library. There is no CPU path; if the shared library is
missing, or no B200 is visible when a compute entry point is called, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# B200_AQP_LIB selects an alternative build of the same library (tuning sweeps only)
LIB_PATH = os.environ.get("B200_AQP_LIB") or os.path.join(PKG_DIR, "libb200aqp.so")

ROW = np.dtype([("key", np.uint32), ("payload", np.uint32)])
TRIPLE = np.dtype([("key", np.uint32), ("Rpayload", np.uint32), ("Spayload", np.uint32)])
TUPLES_PER_CHUNK = (16 * 1024 - 8) // 12
CHUNK_BYTES = 8 + 12 * TUPLES_PER_CHUNK


class Row(C.Structure):
    _fields_ = [("key", C.c_uint32), ("payload", C.c_uint32)]


class Table(C.Structure):          # struct table_t
    _fields_ = [("tuples", C.c_void_p), ("num_tuples", C.c_uint64), ("ratio_holes", C.c_int), ("sorted", C.c_int)]


class ChunkedTable(C.Structure):   # struct chunked_table_t
    _fields_ = [("chunks", C.POINTER(C.c_void_p)), ("current_chunk", C.c_uint64), ("num_chunks", C.c_uint64),
                ("chunk_capacity", C.c_uint64), ("num_tuples", C.c_uint64)]


class Result(C.Structure):         # struct result_t
    _fields_ = [("totalresults", C.c_int64), ("nthreads", C.c_int), ("throughput", C.c_double),
                ("materialized", C.c_int), ("result", C.c_void_p), ("result_type", C.c_int)]


class JoinConfig(C.Structure):     # struct joinconfig_t
    _fields_ = [("NTHREADS", C.c_int), ("PARTFANOUT", C.c_int), ("SCALARSORT", C.c_int), ("SCALARMERGE", C.c_int),
                ("MWAYMERGEBUFFERSIZE", C.c_int), ("NUMASTRATEGY", C.c_int), ("RADIXBITS", C.c_int),
                ("WRITETOFILE", C.c_int), ("MATERIALIZE", C.c_int), ("PRINT", C.c_int), ("CRACKING_THRESHOLD", C.c_int),
                ("ALLOC_CORE", C.c_int)]


class TpchStats(C.Structure):      # struct b200_tpch_stats_t
    _fields_ = [("result_rows", C.c_uint64), ("input_rows", C.c_uint64), ("filtered", C.c_uint64 * 3),
                ("join1_rows", C.c_uint64), ("ms_total", C.c_float), ("ms_filter", C.c_float), ("ms_join", C.c_float),
                ("ms_other", C.c_float), ("kernel_launches", C.c_uint32), ("reserved", C.c_uint32)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k not in ("reserved", "filtered")}
        d["filtered"] = list(self.filtered)
        return d


class LineItemTable(C.Structure):  # TpcHTypes.hpp:50-61
    _fields_ = [("numTuples", C.c_uint64), ("l_orderkey", C.c_void_p), ("l_shipdate", C.c_void_p),
                ("l_commitdate", C.c_void_p), ("l_receiptdate", C.c_void_p), ("l_shipmode", C.c_void_p),
                ("l_partkey", C.c_void_p), ("l_quantity", C.c_void_p), ("l_shipinstruct", C.c_void_p),
                ("l_returnflag", C.c_void_p)]


class OrdersTable(C.Structure):    # TpcHTypes.hpp:63-68
    _fields_ = [("numTuples", C.c_uint64), ("o_orderkey", C.c_void_p), ("o_orderdate", C.c_void_p), ("o_custkey", C.c_void_p)]


class CustomerTable(C.Structure):  # TpcHTypes.hpp:70-75
    _fields_ = [("numTuples", C.c_uint64), ("c_custkey", C.c_void_p), ("c_mktsegment", C.c_void_p),
                ("c_nationkey", C.c_void_p)]


class PartTable(C.Structure):      # TpcHTypes.hpp:77-83
    _fields_ = [("numTuples", C.c_uint64), ("p_partkey", C.c_void_p), ("p_brand", C.c_void_p), ("p_size", C.c_void_p),
                ("p_container", C.c_void_p)]


class JoinStats(C.Structure):      # struct b200_join_stats_t
    _fields_ = [("matches", C.c_int64), ("checksum", C.c_uint64), ("keysum", C.c_uint64), ("radix_bits", C.c_uint32),
                ("num_passes", C.c_uint32), ("bits_pass1", C.c_uint32), ("bits_pass2", C.c_uint32),
                ("kernel_launches", C.c_uint32), ("plan_flags", C.c_uint32), ("ms_total", C.c_float),
                ("ms_hist", C.c_float), ("ms_pass1", C.c_float), ("ms_pass2", C.c_float), ("ms_join", C.c_float),
                ("ms_h2d", C.c_float), ("ms_d2h", C.c_float), ("ms_materialize_host", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


class MgResult(C.Structure):       # struct b200_mg_result_t
    _fields_ = [("matches", C.c_uint64), ("checksum", C.c_uint64), ("keysum", C.c_uint64), ("tuples_sent", C.c_uint64),
                ("tuples_kept", C.c_uint64), ("radix_bits", C.c_uint32), ("bits_pass1", C.c_uint32),
                ("bits_pass2", C.c_uint32), ("world", C.c_uint32), ("kernel_launches", C.c_uint32),
                ("reserved", C.c_uint32), ("ms_total", C.c_float), ("ms_hist", C.c_float), ("ms_scatter", C.c_float),
                ("ms_barrier", C.c_float), ("ms_local", C.c_float), ("ms_reduce", C.c_float), ("ms_pass2", C.c_float),
                ("ms_join", C.c_float)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


# every symbol include/aqp/b200_aqp.h declares: (restype, argtypes)
_vp, _u64, _u32, _u8, _sz, _int = C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint8, C.c_size_t, C.c_int
SYMBOLS = {
    "b200_init": (_int, [_int]),
    "b200_shutdown": (None, []),
    "b200_last_error": (C.c_char_p, []),
    "b200_set_verbose": (None, [_int]),
    "b200_host_alloc": (_vp, [_sz]),
    "b200_host_free": (None, [_vp]),
    "b200_device_alloc": (_vp, [_sz]),
    "b200_device_free": (None, [_vp]),
    "b200_memcpy_h2d": (_int, [_vp, _vp, _sz]),
    "b200_memcpy_d2h": (_int, [_vp, _vp, _sz]),
    "b200_device_sync": (_int, []),
    "run_join": (None, [C.POINTER(Result), C.POINTER(Table), C.POINTER(Table), C.c_char_p, C.POINTER(JoinConfig)]),
    "RHO": (C.POINTER(Result), [C.POINTER(Table), C.POINTER(Table), C.POINTER(JoinConfig)]),
    "destroy_table": (None, [C.POINTER(ChunkedTable)]),
    "b200_last_join_stats": (None, [C.POINTER(JoinStats)]),
    "b200_preload_relations": (_int, [C.POINTER(Table), C.POINTER(Table)]),
    "b200_join_preload": (_int, [C.c_char_p, C.POINTER(JoinConfig), C.POINTER(Result)]),
    "b200_free_preload": (None, []),
    "b200_join_device": (_int, [_vp, _u64, _vp, _u64, _vp, _u64, C.POINTER(JoinStats), _vp]),
    "b200_join_plan": (None, [_u64, C.POINTER(_u32), C.POINTER(_u32), C.POINTER(_u32)]),
    "b200_radix_hist_device": (_int, [_vp, _u64, _u32, _u32, _vp, _vp]),
    "b200_exclusive_scan_u32_device": (_int, [_vp, _u32, _vp, _vp]),
    "b200_radix_scatter_device": (_int, [_vp, _u64, _u32, _u32, _vp, _vp, _vp, _vp]),
    "b200_shard_pass1_device": (_int, [_vp, _u64, _u32, _u32, _u32, _vp, _vp, _vp, _vp]),
    "b200_shard_join_device": (_int, [_vp, _u64, _vp, _vp, _u64, _vp, _vp, _u32, _u32, _u32, _u32, _vp, _vp, _u32,
                                      C.POINTER(JoinStats), _vp]),
    "b200_shard_hist_device": (_int, [_vp, _u64, _u32, _u32, _u32, _vp, _vp, _int, _vp]),
    "b200_shard_scatter_device": (_int, [_vp, _u64, _vp, C.POINTER(_vp), _int, _vp]),
    "b200_copy_async": (_int, [_vp, _vp, _sz, _vp]),
    "b200_exchange_plan_device": (_int, [_vp, _u32, _u32, _u32, _u32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200_shard_join_async_device": (_int, [_vp, _u64, _vp, _vp, _u64, _vp, _vp, _u32, _u32, _u32, _u32, _vp, _vp, _u32,
                                            _vp, _vp]),
    "b200_shard_join_times": (_int, [C.POINTER(JoinStats)]),
    "b200_ipc_export": (_int, [_vp, _vp]),
    "b200_ipc_open": (_int, [_vp, C.POINTER(_vp)]),
    "b200_ipc_close": (_int, [_vp]),
    "b200_mg_unique_id": (_int, [_vp]),
    "b200_mg_init": (_int, [_int, _int, _vp, _u64, _u64]),
    "b200_mg_init_caps": (_int, [_int, _int, _vp, _u64, _u64, _u64, _u32]),
    "b200_mg_join": (_int, [_vp, _u64, _vp, _u64, C.POINTER(MgResult)]),
    "b200_mg_join_materialize": (_int, [_vp, _u64, _vp, _u64, C.POINTER(_vp), C.POINTER(_u64), C.POINTER(MgResult)]),
    "b200_mg_allreduce_u64": (_int, [C.POINTER(_u64), _int]),
    "b200_mg_finalize": (_int, []),
    "seed_generator": (None, [C.c_uint]),
    "create_relation_pk": (_int, [C.POINTER(Table), _u64, _int]),
    "create_relation_fk": (_int, [C.POINTER(Table), _u64, C.c_int64, _int]),
    "create_relation_fk_sel": (_int, [C.POINTER(Table), _u64, C.c_int64, _int]),
    "create_relation_zipf": (_int, [C.POINTER(Table), _u64, C.c_int64, C.c_double, _int]),
    "delete_relation": (None, [C.POINTER(Table)]),
    "b200_gen_pk_device": (_int, [_vp, _u64, _u64, _u64, _u64, _vp]),
    "b200_gen_fk_device": (_int, [_vp, _u64, _u64, _u64, _u64, _u64, _vp]),
    "b200_gen_zipf_device": (_int, [_vp, _u64, C.c_double, _u64, _u64, _u64, _vp]),
    "b200_set_rowid_payload_device": (_int, [_vp, _u64, _u64, _vp]),
    "b200_bitvector_scan_user": (None, [_u8, _u8, _vp, _sz, _vp, C.POINTER(_u64), _sz, _sz, _int]),
    "b200_index_scan_user": (None, [_u8, _u8, _vp, _sz, _vp, _sz, C.POINTER(_sz), C.POINTER(_u64), _sz, _sz, _int]),
    "b200_scan_last_copy_ns": (_u64, []),
    "b200_bitvector_scan_device": (_int, [_u8, _u8, _vp, _sz, _vp, _vp]),
    "b200_scan_count_device": (_int, [_u8, _u8, _vp, _sz, _vp, _vp]),
    "b200_index_scan_device": (_int, [_u8, _u8, _vp, _sz, _u64, _vp, _u64, _vp, _vp]),
    "b200_scan_sum_device": (_int, [_u8, _u8, _vp, _sz, _vp, _vp]),
    "b200_value_scan_device": (_int, [_u8, _u8, _vp, _sz, _vp, _u64, _vp, _vp]),
    "b200_dict_scan_8bit_64bit_device": (_int, [C.c_int64, C.c_int64, _vp, _vp, _sz, _vp, _u64, _vp, _vp]),
    "b200_sum": (_u64, [_u8, _u8, _vp, _sz]),
    "b200_scan": (_u64, [_u8, _u8, _vp, _sz, _vp, _sz]),
    "b200_dict_scan_8bit_64bit": (_u64, [C.c_int64, C.c_int64, _vp, _vp, _sz, _vp, _sz]),
    "b200_explicit_index_scan": (_u64, [_u8, _u8, _vp, _vp, _sz, _vp, _sz]),
    "b200_explicit_index_scan_device": (_int, [_u8, _u8, _vp, _vp, _sz, _vp, _u64, _vp, _vp]),
    "b200_scalar_index_scan": (_u64, [_u8, _u8, _vp, _sz, _vp, _sz]),
    "b200_dict_scan_16bit_64bit": (_u64, [C.c_int64, C.c_int64, _vp, _vp, _sz, _vp, _sz]),
    "b200_dict_scan_32bit_64bit": (_u64, [C.c_int64, C.c_int64, _vp, _sz, _vp, _sz, _vp, _sz]),
    "b200_dict_scan_wide_device": (_int, [_int, _u32, _u32, _vp, _vp, _sz, _vp, _u64, _vp, _vp]),
    "b200_fill_tiled_column_device": (_int, [_vp, _sz, _u64, _vp]),
    "b200_fill_skewed_column_device": (_int, [_vp, _sz, _u64, _u32, _u64, _vp]),
    "b200_kernel_launch_count": (_u64, []),
    # include/aqp/b200_tpch.h
    "tpch_q3": (None, [C.POINTER(Result), C.POINTER(CustomerTable), C.POINTER(OrdersTable), C.POINTER(LineItemTable),
                       C.c_char_p, C.POINTER(JoinConfig)]),
    "tpch_q12": (None, [C.POINTER(Result), C.POINTER(LineItemTable), C.POINTER(OrdersTable), C.c_char_p,
                        C.POINTER(JoinConfig)]),
    "tpch_q19": (None, [C.POINTER(Result), C.POINTER(LineItemTable), C.POINTER(PartTable), C.c_char_p,
                        C.POINTER(JoinConfig)]),
    "b200_tpch_generate_device": (_int, [C.c_double, _u64]),
    "b200_tpch_upload": (_int, [C.POINTER(LineItemTable), C.POINTER(OrdersTable), C.POINTER(CustomerTable),
                                C.POINTER(PartTable)]),
    "b200_tpch_download": (_int, [C.POINTER(LineItemTable), C.POINTER(OrdersTable), C.POINTER(CustomerTable),
                                  C.POINTER(PartTable)]),
    "b200_tpch_free_host": (None, [C.POINTER(LineItemTable), C.POINTER(OrdersTable), C.POINTER(CustomerTable),
                                   C.POINTER(PartTable)]),
    "b200_tpch_free_device": (None, []),
    "b200_tpch_generate_shard_device": (_int, [C.c_double, _u64, _u32, _u32]),
    "b200_tpch_mg_init": (_int, [_int, _int, _vp]),
    "b200_tpch_q12_mg": (_int, [C.POINTER(TpchStats)]),
    "b200_tpch_q3_mg": (_int, [C.POINTER(TpchStats)]),
    "b200_tpch_q19_mg": (_int, [C.POINTER(TpchStats)]),
    "b200_tpch_read_binary": (_int, [C.c_char_p, _int, C.POINTER(LineItemTable), C.POINTER(OrdersTable),
                                     C.POINTER(CustomerTable), C.POINTER(PartTable)]),
    "b200_tpch_write_binary": (_int, [C.c_char_p, _int, C.POINTER(LineItemTable), C.POINTER(OrdersTable),
                                      C.POINTER(CustomerTable), C.POINTER(PartTable)]),
    "b200_tpch_q3_device": (_int, [C.POINTER(TpchStats)]),
    "b200_tpch_q12_device": (_int, [C.POINTER(TpchStats)]),
    "b200_tpch_q19_device": (_int, [C.POINTER(TpchStats)]),
}

_lib = None


def lib():
    """Load libb200aqp.so (fails loudly when it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: build it with `make -C {PKG_DIR}` "
                               "(or __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(L, name)   # AttributeError if the library does not export a declared symbol
            f.restype = res
            f.argtypes = args
        _lib = L
    return _lib


class AqpError(RuntimeError):
    pass


def _st(stream):
    """None -> the library's own stream (NULL in the C ABI); an integer cudaStream_t handle otherwise.
    Handle 0 (what torch reports for its default stream) is passed as cudaStreamLegacy (0x1) because
    NULL already means "library stream" in this ABI."""
    if stream is None:
        return None
    return int(stream) if int(stream) != 0 else 1


def _check(rc, what):
    if rc != 0:
        raise AqpError(f"{what}: {lib().b200_last_error().decode()}")


def init(device: int = -1):
    _check(lib().b200_init(device), "b200_init")


def _table(rel: np.ndarray) -> Table:
    assert rel.dtype == ROW and rel.flags["C_CONTIGUOUS"]
    return Table(rel.ctypes.data, rel.shape[0], 0, 0)


def last_join_stats() -> dict:
    s = JoinStats()
    lib().b200_last_join_stats(C.byref(s))
    return s.as_dict()


def _collect_result(res: Result, materialize: bool, keep_triples: bool = True) -> dict:
    out = {"matches": int(res.totalresults), "nthreads": res.nthreads, "materialized": res.materialized,
           "result_type": res.result_type, "throughput": res.throughput}
    ct = C.cast(res.result, C.POINTER(ChunkedTable))
    if materialize:
        t = ct.contents
        parts = []
        for c in range(t.num_chunks if keep_triples else 0):
            base = t.chunks[c]
            n = C.cast(base, C.POINTER(C.c_uint64))[0]
            assert n <= TUPLES_PER_CHUNK
            buf = (C.c_uint8 * (12 * n)).from_address(base + 8)
            parts.append(np.frombuffer(buf, dtype=TRIPLE).copy())
        if keep_triples:
            out["triples"] = np.concatenate(parts) if parts else np.zeros(0, dtype=TRIPLE)
        out["num_chunks"] = int(t.num_chunks)
        out["table_num_tuples"] = int(t.num_tuples)
    # release exactly as a reference caller does: destroy_table() + free() (tpch.cpp:82)
    lib().destroy_table(ct)
    C.CDLL(None).free(C.c_void_p(res.result))
    out.update({k: v for k, v in last_join_stats().items() if k != "matches"})
    return out


def run_join(R: np.ndarray, S: np.ndarray, materialize: bool = False, nthreads: int = 1, algorithm: bytes = b"RHO",
             keep_triples: bool = True):
    """run_join(result_t*, R, S, "RHO", joinconfig_t*) on HOST relations (joins.hpp:4-6). keep_triples=False leaves
    the materialised chunks uncopied (timing runs: the chunked table is produced, inspected and destroyed)."""
    cfg = JoinConfig()
    cfg.NTHREADS = nthreads
    cfg.MATERIALIZE = int(materialize)
    res = Result()
    tR, tS = _table(R), _table(S)
    lib().run_join(C.byref(res), C.byref(tR), C.byref(tS), algorithm, C.byref(cfg))
    return _collect_result(res, materialize, keep_triples)


def preload_relations(R: np.ndarray, S: np.ndarray):
    tR, tS = _table(R), _table(S)
    _check(lib().b200_preload_relations(C.byref(tR), C.byref(tS)), "b200_preload_relations")


def join_preload(materialize: bool = False, nthreads: int = 1):
    cfg = JoinConfig()
    cfg.NTHREADS = nthreads
    cfg.MATERIALIZE = int(materialize)
    res = Result()
    _check(lib().b200_join_preload(b"RHO", C.byref(cfg), C.byref(res)), "b200_join_preload")
    return _collect_result(res, materialize)


def join_device(d_R: int, nR: int, d_S: int, nS: int, d_out: int = 0, out_capacity: int = 0, stream=None) -> dict:
    s = JoinStats()
    _check(lib().b200_join_device(d_R, nR, d_S, nS, d_out or None, out_capacity, C.byref(s), _st(stream)),
           "b200_join_device")
    return s.as_dict()


def _prefer_bundled_nccl():
    """The library dlopens NCCL on first use. In a Python process torch may be imported later and needs ITS bundled
    libnccl.so.2 (one copy per soname can be loaded): point the library at that copy unless the caller chose one."""
    if "B200_AQP_NCCL_LIB" in os.environ:
        return
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec and spec.submodule_search_locations:
            cand = os.path.join(list(spec.submodule_search_locations)[0], "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["B200_AQP_NCCL_LIB"] = cand
    except Exception:
        pass


def mg_unique_id() -> bytes:
    """128-byte NCCL unique id (rank 0 creates it, every rank passes it to mg_init)"""
    _prefer_bundled_nccl()
    buf = (C.c_ubyte * 128)()
    _check(lib().b200_mg_unique_id(buf), "b200_mg_unique_id")
    return bytes(buf)


def mg_init(rank: int, world: int, unique_id: bytes, nR_total: int, nS_total: int):
    _prefer_bundled_nccl()
    buf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
    _check(lib().b200_mg_init(rank, world, buf, nR_total, nS_total), "b200_mg_init")


def mg_join(d_R: int, nR: int, d_S: int, nS: int) -> dict:
    """The C multi-GPU host's sharded join on this rank's device-resident shard (collective call)."""
    r = MgResult()
    _check(lib().b200_mg_join(d_R, nR, d_S, nS, C.byref(r)), "b200_mg_join")
    return r.as_dict()


def mg_join_materialize(d_R: int, nR: int, d_S: int, nS: int) -> dict:
    """Materialising form: adds d_triples (device address of this rank's {key, Rpayload, Spayload} triples, library
    owned, valid until the next mg_* call) and local_rows to the result dict. Collective."""
    r = MgResult()
    ptr, rows = _vp(), _u64()
    _check(lib().b200_mg_join_materialize(d_R, nR, d_S, nS, C.byref(ptr), C.byref(rows), C.byref(r)), "b200_mg_join_materialize")
    d = r.as_dict()
    d["d_triples"] = ptr.value or 0
    d["local_rows"] = rows.value
    return d


def mg_allreduce_u64(values):
    arr = (_u64 * len(values))(*values)
    _check(lib().b200_mg_allreduce_u64(arr, len(values)), "b200_mg_allreduce_u64")
    return list(arr)


def mg_finalize():
    _check(lib().b200_mg_finalize(), "b200_mg_finalize")


def join_plan(nR: int):
    t, a, b = C.c_uint32(), C.c_uint32(), C.c_uint32()
    lib().b200_join_plan(nR, C.byref(t), C.byref(a), C.byref(b))
    return t.value, a.value, b.value


# ---- device memory helpers (used when torch is not the allocator) -----------------------------------
class DeviceBuffer:
    def __init__(self, nbytes: int):
        init()
        self.nbytes = nbytes
        self.ptr = lib().b200_device_alloc(nbytes)
        if not self.ptr:
            raise AqpError(f"device alloc of {nbytes} bytes failed: {lib().b200_last_error().decode()}")

    def upload(self, a: np.ndarray):
        assert a.nbytes <= self.nbytes
        _check(lib().b200_memcpy_h2d(self.ptr, a.ctypes.data, a.nbytes), "h2d")
        return self

    def download(self, dtype, count) -> np.ndarray:
        a = np.empty(count, dtype=dtype)
        assert a.nbytes <= self.nbytes
        _check(lib().b200_memcpy_d2h(a.ctypes.data, self.ptr, a.nbytes), "d2h")
        return a

    def free(self):
        if self.ptr:
            lib().b200_device_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def to_device(a: np.ndarray) -> DeviceBuffer:
    return DeviceBuffer(max(a.nbytes, 16)).upload(np.ascontiguousarray(a))


# ---- generators ------------------------------------------------------------------------------------------
def _take_relation(t: Table) -> np.ndarray:
    n = t.num_tuples
    buf = (C.c_uint8 * (8 * n)).from_address(t.tuples) if n else b""
    a = np.frombuffer(buf, dtype=ROW).copy()
    lib().delete_relation(C.byref(t))
    return a


def host_gen_pk(n: int, seed: int) -> np.ndarray:
    t = Table()
    lib().seed_generator(seed)
    assert lib().create_relation_pk(C.byref(t), n, 0) == 0
    return _take_relation(t)


def host_gen_fk(n: int, maxid: int, seed: int) -> np.ndarray:
    t = Table()
    lib().seed_generator(seed)
    assert lib().create_relation_fk(C.byref(t), n, maxid, 0) == 0
    return _take_relation(t)


def host_gen_fk_sel(n: int, maxid: int, seed: int) -> np.ndarray:
    t = Table()
    lib().seed_generator(seed)
    assert lib().create_relation_fk_sel(C.byref(t), n, maxid, 0) == 0
    return _take_relation(t)


def host_gen_zipf(n: int, maxid: int, z: float, seed: int) -> np.ndarray:
    t = Table()
    lib().seed_generator(seed)
    assert lib().create_relation_zipf(C.byref(t), n, maxid, z, 0) == 0
    return _take_relation(t)


def gen_pk_device(d_rel: int, n_total: int, seed: int, row_begin: int = 0, n: int | None = None, stream=None):
    _check(lib().b200_gen_pk_device(d_rel, n_total, row_begin, n_total if n is None else n, seed, _st(stream)),
           "b200_gen_pk_device")


def gen_fk_device(d_rel: int, n_total: int, maxid: int, seed: int, row_begin: int = 0, n: int | None = None,
                  stream=None):
    _check(lib().b200_gen_fk_device(d_rel, n_total, maxid, row_begin, n_total if n is None else n, seed,
                                    _st(stream)), "b200_gen_fk_device")


def gen_zipf_device(d_rel: int, n: int, maxid: int, z: float, seed: int, row_begin: int = 0, stream=None):
    _check(lib().b200_gen_zipf_device(d_rel, maxid, z, row_begin, n, seed, _st(stream)), "b200_gen_zipf_device")


# ---- scans -------------------------------------------------------------------------------------------------
def bitvector_scan_user(lo: int, hi: int, data: np.ndarray, num_runs: int = 1, warmup_runs: int = 0,
                        unique_data: bool = True):
    assert data.dtype == np.uint8 and data.flags["C_CONTIGUOUS"]
    n = data.shape[0]
    out = np.zeros(n // 64, dtype=np.uint64)
    t = C.c_uint64(0)
    lib().b200_bitvector_scan_user(lo, hi, data.ctypes.data, n, out.ctypes.data, C.byref(t), num_runs, warmup_runs,
                                   int(unique_data))
    return out, int(t.value)


def index_scan_user(lo: int, hi: int, data: np.ndarray, capacity: int | None = None, num_runs: int = 1,
                    warmup_runs: int = 0, unique_data: bool = True):
    assert data.dtype == np.uint8 and data.flags["C_CONTIGUOUS"]
    n = data.shape[0]
    cap = n if capacity is None else capacity
    out = np.zeros(max(cap, 1), dtype=np.uint64)
    cnt = C.c_size_t(0)
    t = C.c_uint64(0)
    lib().b200_index_scan_user(lo, hi, data.ctypes.data, n, out.ctypes.data, cap, C.byref(cnt), C.byref(t), num_runs,
                               warmup_runs, int(unique_data))
    return out[:min(cnt.value, cap)], int(cnt.value), int(t.value)


def scan_sum(lo: int, hi: int, data: np.ndarray) -> int:
    """SIMD512::sum through the host-buffer entry point"""
    assert data.dtype == np.uint8 and data.flags["C_CONTIGUOUS"]
    return int(lib().b200_sum(lo, hi, data.ctypes.data, data.shape[0]))


def value_scan(lo: int, hi: int, data: np.ndarray, capacity: int | None = None):
    """SIMD512::scan: (matching values as uint32, exact count)"""
    assert data.dtype == np.uint8 and data.flags["C_CONTIGUOUS"]
    cap = data.shape[0] if capacity is None else capacity
    out = np.zeros(max(cap, 1), dtype=np.uint32)
    cnt = int(lib().b200_scan(lo, hi, data.ctypes.data, data.shape[0], out.ctypes.data, cap))
    return out[:min(cnt, cap)], cnt


def dict_scan_8bit_64bit(lo: int, hi: int, dictionary: np.ndarray, data: np.ndarray, capacity: int | None = None):
    """SIMD512::dict_scan_8bit_64bit: (dict[code] of every code whose value is in [lo, hi] as int64, exact count)"""
    assert data.dtype == np.uint8 and data.flags["C_CONTIGUOUS"]
    d = np.ascontiguousarray(dictionary, dtype=np.int64)
    assert d.shape[0] == 256
    cap = data.shape[0] if capacity is None else capacity
    out = np.zeros(max(cap, 1), dtype=np.int64)
    cnt = int(lib().b200_dict_scan_8bit_64bit(lo, hi, d.ctypes.data, data.ctypes.data, data.shape[0], out.ctypes.data, cap))
    return out[:min(cnt, cap)], cnt


def explicit_index_scan(lo: int, hi: int, index: np.ndarray, data: np.ndarray, capacity: int | None = None):
    """SIMD512::explicit_index_scan: (the index entries of the matches, exact count)"""
    assert data.dtype == np.uint8 and index.dtype == np.uint64 and index.shape[0] >= (data.shape[0] // 64 + 7) * 8
    cap = data.shape[0] if capacity is None else capacity
    out = np.zeros(max(cap, 1), dtype=np.uint64)
    cnt = int(lib().b200_explicit_index_scan(lo, hi, index.ctypes.data, data.ctypes.data, data.shape[0], out.ctypes.data, cap))
    return out[:min(cnt, cap)], cnt


def scalar_index_scan(lo: int, hi: int, data: np.ndarray, capacity: int | None = None):
    """scalar_implicit_index_scan: row ids over ALL n values (not only whole 64-value blocks)"""
    assert data.dtype == np.uint8 and data.flags["C_CONTIGUOUS"]
    cap = data.shape[0] if capacity is None else capacity
    out = np.zeros(max(cap, 1), dtype=np.uint64)
    cnt = int(lib().b200_scalar_index_scan(lo, hi, data.ctypes.data, data.shape[0], out.ctypes.data, cap))
    return out[:min(cnt, cap)], cnt


def dict_scan_wide(bits: int, lo: int, hi: int, dictionary: np.ndarray, data: np.ndarray, capacity: int | None = None):
    """SIMD512::dict_scan_16bit_64bit / dict_scan_32bit_64bit: (dict[code] of every selected code, exact count)"""
    d = np.ascontiguousarray(dictionary, dtype=np.int64)
    cap = data.shape[0] if capacity is None else capacity
    out = np.zeros(max(cap, 1), dtype=np.int64)
    if bits == 16:
        assert data.dtype == np.uint16 and d.shape[0] == 1 << 16
        cnt = int(lib().b200_dict_scan_16bit_64bit(lo, hi, d.ctypes.data, data.ctypes.data, data.shape[0], out.ctypes.data, cap))
    else:
        assert bits == 32 and data.dtype == np.uint32
        cnt = int(lib().b200_dict_scan_32bit_64bit(lo, hi, d.ctypes.data, d.shape[0], data.ctypes.data, data.shape[0],
                                                   out.ctypes.data, cap))
    return out[:min(cnt, cap)], cnt


def bitvector_scan_device(lo, hi, d_data, n, d_out, stream=None):
    _check(lib().b200_bitvector_scan_device(lo, hi, d_data, n, d_out, _st(stream)), "b200_bitvector_scan_device")


def scan_count_device(lo, hi, d_data, n, d_count, stream=None):
    _check(lib().b200_scan_count_device(lo, hi, d_data, n, d_count, _st(stream)), "b200_scan_count_device")


def index_scan_device(lo, hi, d_data, n, d_out, capacity, d_count, id_base=0, stream=None):
    _check(lib().b200_index_scan_device(lo, hi, d_data, n, id_base, d_out, capacity, d_count, _st(stream)),
           "b200_index_scan_device")


def kernel_launch_count() -> int:
    return int(lib().b200_kernel_launch_count())


# ---- TPC-H-style pipelines (include/aqp/b200_tpch.h) ---------------------------------------------------------
_TPCH_COLS = {
    "lineitem": (LineItemTable, [("l_orderkey", ROW), ("l_shipdate", np.uint64), ("l_commitdate", np.uint64),
                                 ("l_receiptdate", np.uint64), ("l_shipmode", np.uint8), ("l_partkey", np.uint32),
                                 ("l_quantity", np.float32), ("l_shipinstruct", np.uint8), ("l_returnflag", np.int8)]),
    "orders": (OrdersTable, [("o_orderkey", ROW), ("o_orderdate", np.uint64), ("o_custkey", np.uint32)]),
    "customer": (CustomerTable, [("c_custkey", ROW), ("c_mktsegment", np.uint8), ("c_nationkey", np.uint32)]),
    "part": (PartTable, [("p_partkey", ROW), ("p_brand", np.uint8), ("p_size", np.uint32), ("p_container", np.uint8)]),
}


def tpch_generate_device(scale_factor: float, seed: int = 1):
    init()
    _check(lib().b200_tpch_generate_device(scale_factor, seed), "b200_tpch_generate_device")


def tpch_generate_shard_device(scale_factor: float, seed: int, rank: int, world: int):
    init()
    _check(lib().b200_tpch_generate_shard_device(scale_factor, seed, rank, world), "b200_tpch_generate_shard_device")


def tpch_mg_init(rank: int, world: int, unique_id: bytes):
    _prefer_bundled_nccl()
    buf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
    _check(lib().b200_tpch_mg_init(rank, world, buf), "b200_tpch_mg_init")


def tpch_q12_mg() -> dict:
    s = TpchStats()
    _check(lib().b200_tpch_q12_mg(C.byref(s)), "b200_tpch_q12_mg")
    return s.as_dict()


def tpch_q3_mg() -> dict:
    s = TpchStats()
    _check(lib().b200_tpch_q3_mg(C.byref(s)), "b200_tpch_q3_mg")
    return s.as_dict()


def tpch_q19_mg() -> dict:
    s = TpchStats()
    _check(lib().b200_tpch_q19_mg(C.byref(s)), "b200_tpch_q19_mg")
    return s.as_dict()


def tpch_download() -> dict:
    """Device tables -> dict of dicts of numpy columns (copies)."""
    st = {k: cls() for k, (cls, _) in _TPCH_COLS.items()}
    _check(lib().b200_tpch_download(C.byref(st["lineitem"]), C.byref(st["orders"]), C.byref(st["customer"]),
                                    C.byref(st["part"])), "b200_tpch_download")
    out = {}
    for name, (cls, cols) in _TPCH_COLS.items():
        n = st[name].numTuples
        out[name] = {}
        for col, dt in cols:
            dt = np.dtype(dt)
            buf = (C.c_uint8 * (n * dt.itemsize)).from_address(getattr(st[name], col)) if n else b""
            out[name][col] = np.frombuffer(buf, dtype=dt).copy()
    lib().b200_tpch_free_host(C.byref(st["lineitem"]), C.byref(st["orders"]), C.byref(st["customer"]), C.byref(st["part"]))
    return out


def tpch_write_binary(root: str, scale: int, tables: dict):
    """Host tables -> the reference's binary column files (<root>/scaleNNN/<table>.tbl.dir/...). Host code only."""
    st = {k: tpch_host_struct(k, v) for k, v in tables.items()}
    g = lambda k: C.byref(st[k]) if k in st else None
    _check(lib().b200_tpch_write_binary(root.encode(), scale, g("lineitem"), g("orders"), g("customer"), g("part")),
           "b200_tpch_write_binary")


def tpch_read_binary(root: str, scale: int) -> dict:
    """The reference's binary column files -> dict of dicts of numpy columns (only the columns whose files exist)."""
    st = {k: cls() for k, (cls, _) in _TPCH_COLS.items()}
    _check(lib().b200_tpch_read_binary(root.encode(), scale, C.byref(st["lineitem"]), C.byref(st["orders"]),
                                       C.byref(st["customer"]), C.byref(st["part"])), "b200_tpch_read_binary")
    out = {}
    for name, (cls, cols) in _TPCH_COLS.items():
        n = st[name].numTuples
        out[name] = {}
        for col, dt in cols:
            dt = np.dtype(dt)
            ptr = getattr(st[name], col)
            if ptr:
                buf = (C.c_uint8 * (n * dt.itemsize)).from_address(ptr) if n else b""
                out[name][col] = np.frombuffer(buf, dtype=dt).copy()
    lib().b200_tpch_free_host(C.byref(st["lineitem"]), C.byref(st["orders"]), C.byref(st["customer"]), C.byref(st["part"]))
    return out


def tpch_host_struct(name: str, cols: dict):
    cls, spec = _TPCH_COLS[name]
    st = cls()
    st._keep = []
    n = None
    for col, dt in spec:
        a = np.ascontiguousarray(cols[col])
        assert a.dtype == np.dtype(dt), (col, a.dtype)
        n = len(a) if n is None else n
        st._keep.append(a)
        setattr(st, col, a.ctypes.data)
    st.numTuples = n
    return st


def tpch_upload(tables: dict):
    st = {k: tpch_host_struct(k, v) for k, v in tables.items()}
    g = lambda k: C.byref(st[k]) if k in st else None
    _check(lib().b200_tpch_upload(g("lineitem"), g("orders"), g("customer"), g("part")), "b200_tpch_upload")


def tpch_query_device(q: int) -> dict:
    s = TpchStats()
    f = {3: lib().b200_tpch_q3_device, 12: lib().b200_tpch_q12_device, 19: lib().b200_tpch_q19_device}[q]
    _check(f(C.byref(s)), f"b200_tpch_q{q}_device")
    return s.as_dict()


def tpch_query_host(q: int, tables: dict, nthreads: int = 1) -> dict:
    """The drop-in host entry points tpch_q3 / tpch_q12 / tpch_q19 (tpch.hpp:7-21) on host tables."""
    st = {k: tpch_host_struct(k, v) for k, v in tables.items()}
    res, cfg = Result(), JoinConfig()
    cfg.NTHREADS = nthreads
    if q == 3:
        lib().tpch_q3(C.byref(res), C.byref(st["customer"]), C.byref(st["orders"]), C.byref(st["lineitem"]), b"RHO", C.byref(cfg))
    elif q == 12:
        lib().tpch_q12(C.byref(res), C.byref(st["lineitem"]), C.byref(st["orders"]), b"RHO", C.byref(cfg))
    elif q == 19:
        lib().tpch_q19(C.byref(res), C.byref(st["lineitem"]), C.byref(st["part"]), b"RHO", C.byref(cfg))
    else:
        raise ValueError(q)
    out = {"result_rows": int(res.totalresults), "result_type": res.result_type, "throughput": res.throughput}
    ct = C.cast(res.result, C.POINTER(ChunkedTable))
    lib().destroy_table(ct)
    C.CDLL(None).free(C.c_void_p(res.result))
    return out
