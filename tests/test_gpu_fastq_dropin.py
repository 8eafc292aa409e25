"""The FASTQ ingest under the reference's own loader on the GPU box: oracle/_ref/fastq_dropin_test links
the UNMODIFIED src/ReadData.cpp (+ ReadFilter.cpp, BBHashMap.cpp, dnaToBits.cpp; boost::iostreams
replaced by the zlib stand-in in oracle/ref/boost_shim) with the adaptor
nanospring_b200/cpp/GpuMinHashReadFilter.h.  The same file is loaded by ReadData::loadFromFile(...,
low_mem = true) on the host and by GpuMinHashReadFilter::initializeFromFile on the device; reads,
sketches and online queries must agree.  Built by `make -C oracle fastq_dropin` in the build container
(needs /root/reference); the binary travels to the GPU box."""
import gzip
import os
import subprocess

import numpy as np
import pytest

from fastq_cases import random_fastq
from test_fastq_oracle import fastq_text_of

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "fastq_dropin_test")


def run(path, gz, k, n, thr, tmp):
    r = subprocess.run([BIN, str(path), str(int(gz)), str(k), str(n), str(thr), str(tmp)], capture_output=True,
                       text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="8"))
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert "FASTQ DROPIN OK" in r.stdout and "mismatches 0" in r.stdout
    return r.stdout


@pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/fastq_dropin_test not built (needs /root/reference)")
def test_ci_file_through_both_loaders(tmp_path, c1_reads):
    bases, offsets = c1_reads
    p = tmp_path / "test_file.fastq.gz"
    with gzip.open(p, "wb", compresslevel=1) as f:
        f.write(fastq_text_of(bases, offsets))
    out = run(p, True, 23, 60, 6, tmp_path)
    assert "reads 25004 bases 40747441" in out


@pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/fastq_dropin_test not built (needs /root/reference)")
@pytest.mark.parametrize("end,crlf,gz", [("\n", False, False), ("", False, True), ("\n@tail", True, False),
                                         ("\n@t\nACGT", False, True)])
def test_random_files_through_both_loaders(tmp_path, end, crlf, gz):
    rng = np.random.default_rng(len(end) * 7 + crlf * 3 + gz)
    text = random_fastq(rng, 1500, 2500, crlf=crlf, end=end)
    p = tmp_path / ("r.fastq.gz" if gz else "r.fastq")
    if gz:
        with gzip.open(p, "wb", compresslevel=1) as f:
            f.write(text)
    else:
        p.write_bytes(text)
    run(p, gz, 15, 30, 3, tmp_path)
