"""CPU tests of the drop-in boundary: libnsmh.so loads without a GPU, exports every symbol
include/nsmh.h declares, refuses to compute without a CUDA device (no CPU fallback), and its
host-only helpers agree with the oracle."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import nanospring_b200 as ns
from nanospring_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        return False


def test_library_exports_every_declared_symbol():
    L = ns.lib()
    names = _lib.exported_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(L, name), f"{name} declared in include/nsmh.h but not exported"
    # and the binding table covers the header
    assert set(_lib._SIGS) | {"nsmh_last_error", "nsmh_version"} == set(names)


def test_header_compiles_as_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "nsmh.h"\nint main(void){ nsmh_stats s; (void)s; return NSMH_OK; }\n')
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-c", str(src), "-o", str(tmp_path / "t.o")])


def test_struct_layouts_match_the_python_bindings(tmp_path):
    """sizeof / offsetof of nsmh_stats and nsmh_synth_params as the C compiler sees them vs the ctypes mirrors."""
    fields = [f for f, _ in _lib.Stats._fields_]
    src = tmp_path / "layout.c"
    body = "".join(f'printf("{f} %zu\\n", offsetof(nsmh_stats, {f}));\n' for f in fields)
    src.write_text('#include <stddef.h>\n#include <stdio.h>\n#include "nsmh.h"\nint main(void){\n'
                   'printf("sizeof %zu\\n", sizeof(nsmh_stats));\n' + body + 'return 0; }\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["/usr/bin/gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    assert int(out["sizeof"]) == C.sizeof(_lib.Stats)
    for f in fields:
        assert int(out[f]) == getattr(_lib.Stats, f).offset, f


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_rand_from_seed_matches_oracle(orc):
    for seed, n in [(20261017, 60), (1, 30), (7, 120), (0, 1), (0xFFFFFFFF, 400)]:
        assert (ns.rand_from_seed(seed, n) == orc.rand_from_seed(seed, n)).all()


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    f = ns.MinHashReadFilter()
    f.randNumbers = ns.rand_from_seed(1, 60)
    with pytest.raises(ns.NsmhError) as ei:
        f.initialize([b"ACGTACGTACGTACGTACGTACGTACGT"])
    assert ei.value.code == _lib.NSMH_ECUDA
    assert "no CUDA device" in str(ei.value)


def test_argument_validation_happens_before_cuda():
    h = C.c_void_p()
    rnd = ns.rand_from_seed(1, 4)
    p = rnd.ctypes.data_as(_lib.u64p)
    L = ns.lib()
    assert L.nsmh_create(0, 4, 1, p, 0, C.byref(h)) == _lib.NSMH_EINVAL      # k == 0 (main.cpp:120)
    assert L.nsmh_create(32, 4, 1, p, 0, C.byref(h)) == _lib.NSMH_EINVAL     # k == 32 is UB upstream
    assert L.nsmh_create(23, 0, 1, p, 0, C.byref(h)) == _lib.NSMH_EINVAL
    assert L.nsmh_create(23, 4, 1, None, 0, C.byref(h)) == _lib.NSMH_EINVAL
    assert b"rand_numbers" in L.nsmh_last_error()
    assert L.nsmh_sketch(None) == _lib.NSMH_EINVAL
    assert L.nsmh_destroy(None) == _lib.NSMH_OK


def test_synthetic_reads_follow_the_recipe():
    """createData.py semantics: deterministic, subset-consistent, ~10% edits, ~half reverse
    complemented, alphabet ACGT."""
    p = ns.synth_params(genome_len=50_000, genome_seed=1, read_seed=2)
    lengths = ns.synth_lengths(200, 1000, seed=4)
    a = ns.synth_reads_host(lengths, p)
    b = ns.synth_reads_host(lengths, p)
    assert (a.bases == b.bases).all()
    sub = ns.synth_reads_host(lengths[50:60], p, first_read=50)
    for i in range(10):
        assert sub.getRead(i) == a.getRead(50 + i)
    assert set(a.bases.tobytes()) <= set(b"ACGT")
    # error-free forward reads are substrings of the (implicit) genome: two reads with p=0
    clean = ns.synth_params(genome_len=5_000, genome_seed=1, read_seed=2, p_ins=0, p_del=0, p_sub=0, p_rc=0)
    full = ns.synth_reads_host(np.array([20_000], dtype=np.uint64), clean).getRead(0)
    assert full[:5000] == full[5000:10000]             # wrap-around repeats the genome
    rot = full[:5000] + full[:5000]
    other = ns.synth_reads_host(np.array([300, 300, 300], dtype=np.uint64), clean).getRead(2)
    assert other in rot
    # strands: with p_rc = 1 the read is the reverse complement of the p_rc = 0 read
    rc1 = ns.synth_params(genome_len=5_000, genome_seed=1, read_seed=2, p_ins=0, p_del=0, p_sub=0, p_rc=1.0)
    r1 = ns.synth_reads_host(np.array([300, 300, 300], dtype=np.uint64), rc1).getRead(2)
    assert r1 == ns.reverse_complement(other)


def test_synth_lengths_distribution():
    L = ns.synth_lengths(20000, 10000, seed=2)
    assert L.min() >= 1 and abs(L.mean() / 10000 - 1) < 0.05
    assert (ns.synth_lengths(10, 500, dist="const") == 500).all()
