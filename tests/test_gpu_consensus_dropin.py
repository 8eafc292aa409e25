"""The drop-in under the REAL caller: oracle/_ref/consensus_dropin_test links
the reference's own, unmodified consensus stage (Consensus.cpp, ConsensusGraph.cpp, minimap2 ...)
and runs Consensus::generateAndWriteConsensus twice over the same reads - once with the
reference's MinHashReadFilter (CPU), once with GpuMinHashReadFilter (libnsmh.so) behind the same
ReadFilter* - and requires every output stream to be byte-identical; then once more with the GPU
filter queried from 8 OpenMP threads at once (Consensus.cpp:29).  Built by
`make -C oracle consensus_dropin` in the build container; the binary travels to the GPU box."""
import os
import subprocess

import pytest

import nanospring_b200 as ns
from test_gpu_dropin import write_reads

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "consensus_dropin_test")
needs_bin = pytest.mark.skipif(not os.path.exists(BIN),
                               reason="oracle/_ref/consensus_dropin_test not built (needs /root/reference)")


def make_reads(tmp_path, n_reads, k):
    lengths = ns.synth_lengths(n_reads, 3000, seed=21)
    lengths[:4] = [0, 5, k - 1, k]
    rd = ns.synth_reads_host(lengths, ns.synth_params(genome_len=200 * n_reads, genome_seed=3, read_seed=4,
                                                      p_ins=0.01, p_del=0.01, p_sub=0.02))
    p = tmp_path / "reads.bin"
    write_reads(p, rd.bases, rd.offsets)
    return p


@needs_bin
def test_reference_consensus_stage_is_deterministic_single_threaded(tmp_path):
    """CPU only: the premise of the GPU comparison below (one thread => reproducible streams)."""
    p = make_reads(tmp_path, 300, 23)
    outs = []
    for _ in range(2):
        r = subprocess.run([BIN, str(p), "23", "60", "6", str(tmp_path), "ref"], capture_output=True, text=True,
                           timeout=600)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        outs.append(r.stdout.splitlines()[1:])          # line 0 carries the wall time
    assert outs[0] == outs[1] and any("metaData" in line for line in outs[0])


@pytest.mark.gpu
@needs_bin
@pytest.mark.parametrize("k,n,thr", [(23, 60, 6), (15, 30, 3)])
def test_consensus_streams_identical_with_gpu_filter(tmp_path, k, n, thr):
    p = make_reads(tmp_path, 1200, k)
    r = subprocess.run([BIN, str(p), str(k), str(n), str(thr), str(tmp_path), "both", "8"], capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2500:] + r.stderr[-1500:]
    assert "CONSENSUS DROPIN OK" in r.stdout
    print(r.stdout)


@pytest.mark.gpu
@needs_bin
@pytest.mark.parametrize("ndev", [2, 8])
def test_consensus_streams_identical_with_filter_on_several_devices(tmp_path, ndev):
    """The same with GpuMinHashReadFilter::devices set: ONE process, the reference's OpenMP threads, every GPU of
    the box behind the one ReadFilter* (csrc/multidev.cu).  On a box with fewer GPUs the devices are named
    round-robin (two handles on one device exercise the same code)."""
    import torch
    have = max(torch.cuda.device_count(), 1)
    p = make_reads(tmp_path, 1200, 23)
    env = dict(os.environ, DROPIN_DEVICES=",".join(str(i % have) for i in range(ndev)))
    r = subprocess.run([BIN, str(p), "23", "60", "6", str(tmp_path), "both", "8"], capture_output=True, text=True,
                       timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-2500:] + r.stderr[-1500:]
    assert "CONSENSUS DROPIN OK" in r.stdout and f"gpu filter on {ndev} devices" in r.stdout
    print(r.stdout)
