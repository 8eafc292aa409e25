"""The C++ drop-in boundary on the GPU box: oracle/_ref/dropin_test is the adaptor
nanospring_b200/cpp/GpuMinHashReadFilter.h compiled against the REFERENCE's own headers
(ReadFilter.h, ReadData.h) and driven through a ReadFilter* exactly like the reference's
callers (Compressor.cpp:69-76, Consensus.cpp:29,189), with the reference's MinHashReadFilter
(libnsref.so) answering the same queries on the CPU.  Built by `make -C oracle dropin` in the
build container (needs /root/reference); the binary travels to the GPU box."""
import os
import struct
import subprocess

import numpy as np
import pytest

import nanospring_b200 as ns

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "dropin_test")


def write_reads(path, bases, offsets):
    with open(path, "wb") as f:
        f.write(struct.pack("<I", offsets.size - 1))
        f.write(np.ascontiguousarray(offsets, dtype=np.uint64).tobytes())
        f.write(np.ascontiguousarray(bases, dtype=np.uint8).tobytes())


@pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/dropin_test not built (needs /root/reference)")
@pytest.mark.parametrize("which,k,n,thr", [("edge", 23, 60, 6), ("synth", 23, 60, 6), ("synth", 15, 30, 3)])
def test_cpp_adaptor_is_a_drop_in(tmp_path, edge, which, k, n, thr):
    if which == "edge":
        bases, offsets = edge["bases"], edge["offsets"]
    else:
        lengths = ns.synth_lengths(1200, 2500, seed=21)
        lengths[:4] = [0, 5, k - 1, k]
        rd = ns.synth_reads_host(lengths, ns.synth_params(genome_len=150_000, genome_seed=3, read_seed=4,
                                                          p_ins=0.01, p_del=0.01, p_sub=0.02))
        bases, offsets = rd.bases, rd.offsets
    p = tmp_path / "reads.bin"
    write_reads(p, bases, offsets)
    env = dict(os.environ, OMP_NUM_THREADS="8")
    r = subprocess.run([BIN, str(p), str(k), str(n), str(thr), str(tmp_path)], capture_output=True, text=True,
                       timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert "DROPIN OK" in r.stdout and "mismatches 0" in r.stdout


@pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/dropin_test not built (needs /root/reference)")
@pytest.mark.parametrize("mode,ndev", [("ascii", 2), ("bitset", 1), ("bitset", 3)])
def test_cpp_adaptor_multi_device_and_bitset_input(tmp_path, mode, ndev):
    """The same caller pattern with (a) several devices behind the one filter (`devices`: one process, all GPUs,
    the reference's process model) and (b) the reads taken from the reference's 2-bit store (the low-memory temp
    file tempDir/readBitset + the read lengths) instead of getRead()."""
    import torch
    k, n, thr = 23, 60, 6
    lengths = ns.synth_lengths(1500, 2500, seed=33)
    lengths[:5] = [0, 1, 2, k - 1, k]
    rd = ns.synth_reads_host(lengths, ns.synth_params(genome_len=150_000, genome_seed=5, read_seed=6,
                                                      p_ins=0.01, p_del=0.01, p_sub=0.02))
    p = tmp_path / "reads.bin"
    write_reads(p, rd.bases, rd.offsets)
    have = torch.cuda.device_count()
    env = dict(os.environ, OMP_NUM_THREADS="8", DROPIN_MODE=mode)
    if ndev > 1:
        env["DROPIN_DEVICES"] = ",".join(str(i % have) for i in range(ndev))
    r = subprocess.run([BIN, str(p), str(k), str(n), str(thr), str(tmp_path)], capture_output=True, text=True,
                       timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert "DROPIN OK" in r.stdout and "mismatches 0" in r.stdout
