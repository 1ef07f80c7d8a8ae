import hashlib

import numpy as np


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def sorted_triples(t):
    return np.sort(t, order=["key", "Rpayload", "Spayload"])


def expected_pkfk(R, S):
    """Closed-form expectation for a join whose build side has unique keys: numpy, independent of
    both the oracle and the GPU path."""
    kmax = int(max(R["key"].max(initial=0), S["key"].max(initial=0))) + 1
    pay = np.zeros(kmax, dtype=np.uint64)
    present = np.zeros(kmax, dtype=bool)
    pay[R["key"]] = R["payload"]
    present[R["key"]] = True
    hit = present[S["key"]]
    matches = int(hit.sum())
    checksum = int(pay[S["key"][hit]].sum() + S["payload"][hit].astype(np.uint64).sum())
    keysum = int(S["key"][hit].astype(np.uint64).sum())
    return matches, checksum, keysum
