"""SURVEY 8(f) N4, oracle only (no device code exists for this row yet): the UNMODIFIED ConsensusGraph::alignRead
(ConsensusGraph.cpp:161-398 = minimap2 index of the main path, mm_map, CIGAR -> edit script), reached through
oracle/_ref/libnsref_consensus.so (oracle/ref/consensus_harness.cpp: nsref_align_read), pinned to golden vectors
generated from it (tests/golden/make_align_golden.py) - and the parity contract a device implementation would have
to meet, written down as a checkable property: the edit script, applied to the main path from beginOffset,
reproduces the read."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden", "align_golden.npz")
SO = os.path.join(ROOT, "oracle", "_ref", "libnsref_consensus.so")
SAME, INSERT, DELETE, SUBSTITUTION = 0, 1, 2, 3


def gen_module():
    spec = importlib.util.spec_from_file_location("make_align_golden", os.path.join(ROOT, "tests", "golden", "make_align_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def cases():
    z = np.load(GOLDEN)
    for i, (ok, rp, bo, eo) in enumerate(z["meta"].tolist()):
        yield i, z[f"ref{i}"].tobytes(), z[f"read{i}"].tobytes(), ok, rp, bo, eo, z[f"types{i}"], z[f"infos{i}"]


def test_edit_script_applied_to_the_main_path_gives_the_read():
    """The contract, on the golden vectors alone (runs anywhere).  alignRead's comments (ConsensusGraph.cpp:243-268):
    beginOffset > 0 = where the alignment starts on the main path, the read's soft-clipped head is in the script as
    insertions; beginOffset <= 0 = the read starts |beginOffset| bases left of the main path, those bases are NOT
    in the script.  endOffset < 0 = the alignment ends |endOffset| bases before the main path's end (soft-clipped
    tail in the script); endOffset >= 0 = the read's last endOffset bases lie beyond the main path and are not in
    the script.  relPos = reference start - query start of the alignment."""
    aligned = 0
    for i, ref, read, ok, rp, bo, eo, types, infos in cases():
        if not ok:
            assert types.size == 0
            continue
        aligned += 1
        pos = max(bo, 0)
        out = bytearray(read[:max(-bo, 0)])
        for t, v in zip(types.tolist(), infos.tolist()):
            if t == SAME:
                out += ref[pos:pos + v]
                pos += v
            elif t == INSERT:
                out.append(v)
            elif t == DELETE:
                assert ref[pos] == v, f"case {i}: deleted base is not the main path's"
                pos += 1
            else:
                raise AssertionError("alignRead emits no substitutions (ConsensusGraph.cpp:300-390)")
        if eo > 0:
            out += read[len(read) - eo:]
        assert bytes(out) == read, f"case {i}: script does not reproduce the read"
        assert pos == (len(ref) + eo if eo < 0 else len(ref)), f"case {i}: main-path position after the script"
        assert (types == SAME).any()                      # updateGraph's precondition (ConsensusGraph.h:252)
    assert aligned >= 10


@pytest.mark.skipif(not os.path.exists(SO), reason="oracle/_ref/libnsref_consensus.so not built (needs /root/reference)")
def test_reference_alignread_reproduces_the_golden_vectors():
    m = gen_module()
    L = m.load()
    z = np.load(GOLDEN)
    assert z["params"].tolist() == [m.M_K, m.M_W, m.MAX_CHAIN_ITER]
    for i, ref, read, ok, rp, bo, eo, types, infos in cases():
        got = m.align(L, ref, read)
        assert got[:4] == (ok, rp, bo, eo), f"case {i}"
        assert got[4].size == types.size and (got[4] == types).all() and (got[5] == infos).all(), f"case {i}: edit script"
    # the generator's cases are what the file holds
    for i, (ref, read, _) in enumerate(m.cases()):
        assert z[f"ref{i}"].tobytes() == ref and z[f"read{i}"].tobytes() == read
