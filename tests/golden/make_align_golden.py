#!/usr/bin/env python
"""Golden vectors for SURVEY 8(f) N4: the UNMODIFIED ConsensusGraph::alignRead (ConsensusGraph.cpp:161-398, i.e.
minimap2 index of the main path + mm_map + CIGAR -> edit script) run through oracle/_ref/libnsref_consensus.so
(oracle/ref/consensus_harness.cpp: nsref_align_read) on seeded synthetic (main path, read) pairs.

    python tests/golden/make_align_golden.py        # needs /root/reference (oracle/_ref built); writes align_golden.npz

Cases: reads that overlap the middle, the left end and the right end of the main path (negative / positive offsets,
soft clips as insertions), error rates from 0 to 15 %, a read that contains the whole main path, unrelated reads
(alignRead returns false), reads shorter than the minimizer window, low-complexity sequence."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SO = os.path.join(ROOT, "oracle", "_ref", "libnsref_consensus.so")
M_K, M_W, MAX_CHAIN_ITER = 20, 50, 400          # main.cpp:63-68


def load():
    L = C.CDLL(SO)
    L.nsref_align_read.restype = C.c_long
    L.nsref_align_read.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_size_t,
                                   C.POINTER(C.c_int), C.POINTER(C.c_long), C.POINTER(C.c_long), C.POINTER(C.c_long),
                                   C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.c_size_t]
    return L


def align(L, ref, read, m_k=M_K, m_w=M_W, iters=MAX_CHAIN_ITER):
    cap = 2 * (len(ref) + len(read)) + 16
    types = np.zeros(cap, dtype=np.uint8)
    infos = np.zeros(cap, dtype=np.uint64)
    ok, rp, bo, eo = C.c_int(0), C.c_long(0), C.c_long(0), C.c_long(0)
    n = L.nsref_align_read(ref, len(ref), read, len(read), m_k, m_w, iters, C.byref(ok), C.byref(rp), C.byref(bo),
                           C.byref(eo), types.ctypes.data_as(C.POINTER(C.c_uint8)), infos.ctypes.data_as(C.POINTER(C.c_uint64)), cap)
    assert 0 <= n <= cap, n
    return ok.value, rp.value, bo.value, eo.value, types[:n].copy(), infos[:n].copy()


def mutate(rng, s, p_sub, p_ins, p_del):
    out = bytearray()
    for c in s:
        r = rng.random()
        if r < p_del:
            continue
        if r < p_del + p_sub:
            out.append(rng.choice([x for x in b"ACGT" if x != c]))
        else:
            out.append(c)
        if rng.random() < p_ins:
            out.append(rng.choice(list(b"ACGT")))
    return bytes(out)


def cases(seed=20261018):
    rng = np.random.default_rng(seed)
    genome = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), size=60000))
    out = []
    for i, (err, rlen, where) in enumerate([(0.0, 3000, "mid"), (0.02, 5000, "mid"), (0.10, 8000, "mid"), (0.15, 4000, "mid"),
                                            (0.05, 6000, "left"), (0.05, 6000, "right"), (0.10, 9000, "left"), (0.10, 9000, "right"),
                                            (0.03, 14000, "cover"), (0.0, 1200, "mid"), (0.08, 700, "mid")]):
        g0 = 2000 + 4000 * i
        ref = genome[g0:g0 + 10000]
        if where == "mid":
            a = g0 + 2500
        elif where == "left":
            a = g0 - rlen // 2
        elif where == "right":
            a = g0 + 10000 - rlen // 2
        else:
            a = g0 - 2000
        read = mutate(rng, genome[a:a + rlen], err * 0.4, err * 0.3, err * 0.3)
        out.append((ref, read, f"{where} err={err} len={rlen}"))
    ref = genome[1000:9000]
    out.append((ref, bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), size=5000)), "unrelated read"))
    out.append((ref, genome[3000:3040], "read shorter than the minimizer window"))
    out.append((ref, b"", "empty read"))
    out.append((ref, b"AC" * 1500, "low-complexity read"))
    out.append((genome[20000:20300], genome[20050:20250], "short main path"))
    out.append((ref, mutate(rng, genome[2000:8000], 0.01, 0.0, 0.0), "substitutions only"))
    return out


def main():
    L = load()
    blob = {}
    meta = []
    for i, (ref, read, what) in enumerate(cases()):
        ok, rp, bo, eo, types, infos = align(L, ref, read)
        blob[f"ref{i}"] = np.frombuffer(ref, np.uint8)
        blob[f"read{i}"] = np.frombuffer(read, np.uint8)
        blob[f"types{i}"] = types
        blob[f"infos{i}"] = infos
        meta.append((ok, rp, bo, eo))
        print(f"{i:2d} {what:45s} ok={ok} relPos={rp} begin={bo} end={eo} edits={types.size} "
              f"(same runs {int((types == 0).sum())}, ins {int((types == 1).sum())}, del {int((types == 2).sum())}, sub {int((types == 3).sum())})")
    blob["meta"] = np.asarray(meta, dtype=np.int64)
    blob["params"] = np.asarray([M_K, M_W, MAX_CHAIN_ITER], dtype=np.int64)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "align_golden.npz"), **blob)


if __name__ == "__main__":
    main()
