"""Generate tests/golden/*.json|npz from the UNMODIFIED reference compiled into oracle/_ref/.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
The fixtures are committed; tests never read /root/reference.

What is recorded (the reference ships no join/scan KATs of its own):
  * generator: first 8 keys + sha256 of the key column for several (seed, size) pairs of
    create_relation_pk / _fk / _fk_sel                         (generator.cpp:352,:474,:515)
  * join: matches / checksum / keysum / sha256(sorted triples) of RHO() on those relations with
    payload = row id, for both the forced-2-pass and the automatic pass count builds
  * zipf: a small reference-generated Zipf relation is stored verbatim (it is not reproducible,
    genzipf.cpp:44-45) together with the reference join result on it
  * scan: popcounts and sha256 of bitvector / row-id outputs of SIMD512 on the tiled column
  * tpch: row counts of the reference's tpch_q3 / tpch_q12 / tpch_q19 on oracle.synth_tpch(sf, seed) tables
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle as O  # noqa: E402


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def sorted_triples(t):
    return np.sort(t, order=["key", "Rpayload", "Spayload"])


def main():
    assert O.have_ref(), "build oracle/_ref first: make -C oracle ref"
    g = {"generator": [], "join": [], "scan": [], "tpch": []}

    gen_cases = [("pk", 1 << 10, None, 11111), ("pk", 1 << 20, None, 11111), ("pk", 100003, None, 12345),
                 ("pk", 1 << 24, None, 11111),
                 ("fk", 1 << 12, 1 << 10, 22222), ("fk", 1 << 22, 1 << 20, 22222), ("fk", 250007, 100003, 54321),
                 ("fk", 1 << 26, 1 << 24, 22222),
                 ("fk_sel", 1 << 16, 100 * (1 << 14) // 50, 22222), ("fk_sel", 1 << 16, 100 * (1 << 14) // 10, 22222)]
    for kind, n, maxid, seed in gen_cases:
        rel = {"pk": lambda: O.ref_gen_pk(n, seed), "fk": lambda: O.ref_gen_fk(n, maxid, seed),
               "fk_sel": lambda: O.ref_gen_fk_sel(n, maxid, seed)}[kind]()
        g["generator"].append({"kind": kind, "n": n, "maxid": maxid, "seed": seed,
                               "first8": [int(x) for x in rel["key"][:8]], "sha256_keys": sha(rel["key"])})

    join_cases = [(1 << 10, 1 << 12, "fk", None), (1 << 16, 1 << 18, "fk", None), (1 << 20, 1 << 22, "fk", None),
                  (100003, 250007, "fk", None), (1 << 14, 1 << 16, "fk_sel", 50), (1 << 14, 1 << 16, "fk_sel", 10),
                  (5000, 1 << 16, "fk", None)]
    for nR, nS, kind, sel in join_cases:
        R = O.set_rowid_payload(O.ref_gen_pk(nR, 11111))
        if kind == "fk":
            S = O.ref_gen_fk(nS, nR, 22222)
        else:
            S = O.ref_gen_fk_sel(nS, 100 * nR // sel, 22222)
        O.set_rowid_payload(S)
        for force2 in (True, False):
            r = O.ref_rho(R, S, nthreads=4, materialize=True, force_2_passes=force2)
            g["join"].append({"nR": nR, "nS": nS, "kind": kind, "sel": sel, "force_2_passes": force2, "nthreads": 4,
                              "matches": r["matches"], "checksum": r["checksum"], "keysum": r["keysum"],
                              "sha256_sorted_triples": sha(sorted_triples(r["triples"]))})

    # Zipf: store the bytes
    nR, nS = 1 << 12, 1 << 15
    zipf = {}
    R = O.set_rowid_payload(O.ref_gen_pk(nR, 11111))
    for z in (0.5, 1.0, 1.5):
        S = O.set_rowid_payload(O.ref_gen_zipf(nS, nR, z))
        r = O.ref_rho(R, S, nthreads=4, materialize=True)
        zipf[f"S_z{z}"] = S["key"].copy()
        g["join"].append({"nR": nR, "nS": nS, "kind": "zipf", "z": z, "force_2_passes": True, "nthreads": 4,
                          "matches": r["matches"], "checksum": r["checksum"], "keysum": r["keysum"],
                          "sha256_sorted_triples": sha(sorted_triples(r["triples"]))})
    np.savez_compressed(os.path.join(HERE, "zipf_inputs.npz"), **zipf)

    n = 1 << 20
    col = O.aligned_u8(n)
    col[:] = O.tiled_column(n)
    rng = np.random.default_rng(7)
    rnd = O.aligned_u8(n)
    rnd[:] = rng.integers(0, 256, n, dtype=np.uint8)
    for name, data in (("tiled", col), ("random_seed7", rnd)):
        for lo, hi in ((0, 0), (0, 26), (0, 128), (0, 255), (5, 5), (17, 200), (100, 50), (255, 255), (128, 255)):
            bv = O.ref_bitvector_scan(lo, hi, data)
            ids = O.ref_index_scan(lo, hi, data)
            g["scan"].append({"column": name, "n": n, "lo": lo, "hi": hi, "count": O.ref_scan_count(lo, hi, data),
                              "sha256_bitvector": sha(bv), "sha256_rowids": sha(ids)})

    # TPC-H-style pipelines: the reference's tpch_q3 / q12 / q19 (SIMD filters + RHO) on seeded numpy tables
    for sf, seed in ((0.02, 5), (0.05, 3)):
        t = O.synth_tpch(sf, seed)
        entry = {"sf": sf, "seed": seed}
        for q in (3, 12, 19):
            r = O.ref_tpch_query(q, t, nthreads=4)
            entry[f"q{q}"] = {k: v for k, v in r.items() if k != "seconds"}
        g["tpch"].append(entry)

    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(g, f, indent=1)
    print("wrote", len(g["generator"]), "generator,", len(g["join"]), "join,", len(g["scan"]), "scan fixtures")


if __name__ == "__main__":
    main()
