"""GPU parity of the FASTQ ingest (csrc/fastq.cu through the C ABI):
nsmh_load_fastq / _device / _file, nsmh_read_offsets, nsmh_get_reads_ascii, nsmh_set_params against
  * the committed outputs of the UNMODIFIED reference loader (tests/golden/fastq_golden.npz),
  * the CPU oracle on seeded texts,
  * the existing ASCII load path (same packed stream => same sketches and candidates),
and the reference's CI file end to end from a .gz file to the golden sketch / candidate checksums."""
import gzip
import json
import os

import numpy as np
import pytest

import nanospring_b200 as ns
from nanospring_b200 import _lib
from nanospring_b200.filter import GpuReadData, MinHashReadFilter, ReadData

from fastq_cases import random_fastq
from test_fastq_oracle import fastq_text_of, fq_golden  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu


def check_loaded(rd, want_bases, want_offsets, what):
    assert rd.getNumReads() == want_offsets.size - 1, what
    assert (rd.offsets == want_offsets).all(), what
    n = rd.getNumReads()
    got = rd.getReads(0, n) if n else np.zeros(0, np.uint8)
    assert got.size == want_bases.size and (got == want_bases).all(), what
    lens = np.diff(want_offsets.astype(np.int64))
    assert rd.maxReadLen == (int(lens.max()) if lens.size else 0)
    assert rd.avgReadLen == (int(lens.sum() // lens.size) if lens.size else 0)


def test_reference_goldens_host_text(fq_golden):  # noqa: F811
    cases, _ = fq_golden
    rd = GpuReadData()
    for text, bases, offsets in cases:
        rd.loadFromText(text)
        check_loaded(rd, bases, offsets, text[:40])
        # single reads, as ReadData::getRead hands them out
        for i in list(range(min(rd.getNumReads(), 5))) + ([rd.getNumReads() - 1] if rd.getNumReads() else []):
            assert rd.getRead(i) == bases[int(offsets[i]):int(offsets[i + 1])].tobytes()
    rd.close()


def test_reference_goldens_device_text_any_alignment(fq_golden):  # noqa: F811
    import torch
    cases, _ = fq_golden
    rd = GpuReadData()
    for ci, (text, bases, offsets) in enumerate(cases):
        for mis in (0, 1, 4, 7):
            buf = torch.full((len(text) + mis + 1,), 10, dtype=torch.uint8, device="cuda")   # '\n' behind the text
            if text:
                buf[mis:mis + len(text)] = torch.from_numpy(np.frombuffer(text, np.uint8).copy()).cuda()
            torch.cuda.synchronize()
            rd.loadFromDeviceText(buf.data_ptr() + mis, len(text))
            check_loaded(rd, bases, offsets, f"case {ci} misalignment {mis}")
    rd.close()


def test_files_plain_gzip_multimember(fq_golden, tmp_path):  # noqa: F811
    cases, _ = fq_golden
    rd = GpuReadData()
    for ci, (text, bases, offsets) in enumerate(cases):
        p = tmp_path / f"c{ci}.fastq"
        p.write_bytes(text)
        rd.loadFromFile(str(p), GpuReadData.FASTQ)
        check_loaded(rd, bases, offsets, f"case {ci} plain file")
        pz = tmp_path / f"c{ci}.fastq.gz"
        with gzip.open(pz, "wb") as f:
            f.write(text)
        rd.loadFromFile(str(pz), GpuReadData.GZIP, low_mem=True)
        check_loaded(rd, bases, offsets, f"case {ci} gzip file")
        if len(text) > 10:                     # the same text as two concatenated gzip members
            cut = len(text) // 3
            pm = tmp_path / f"c{ci}.mm.gz"
            pm.write_bytes(gzip.compress(text[:cut]) + gzip.compress(text[cut:]))
            rd.loadFromFile(str(pm), GpuReadData.GZIP)
            check_loaded(rd, bases, offsets, f"case {ci} two gzip members")
    rd.close()


def test_file_errors(tmp_path):
    rd = GpuReadData()
    with pytest.raises(ns.NsmhError) as ei:
        rd.loadFromFile(str(tmp_path / "missing.fastq"), GpuReadData.FASTQ)
    assert ei.value.code == _lib.NSMH_EINVAL
    good = gzip.compress(b"@a\nACGT\n+\nIIII\n" * 1000)
    (tmp_path / "trunc.gz").write_bytes(good[:len(good) // 2])
    with pytest.raises(ns.NsmhError):
        rd.loadFromFile(str(tmp_path / "trunc.gz"), GpuReadData.GZIP)
    (tmp_path / "junk.gz").write_bytes(b"this is not gzip data at all" * 10)
    with pytest.raises(ns.NsmhError):
        rd.loadFromFile(str(tmp_path / "junk.gz"), GpuReadData.GZIP)
    with pytest.raises(ValueError):
        rd.loadFromFile(str(tmp_path / "x.reads"), GpuReadData.READ)
    # the handle is still usable afterwards
    rd.loadFromText(b"@a\nACGT\n+\nIIII\n")
    assert rd.getNumReads() == 1 and rd.getRead(0) == b"ACGT"
    f = MinHashReadFilter()
    with pytest.raises(RuntimeError):
        f.initialize(GpuReadData())
    rd.close()


def test_oracle_random_texts(orc):
    rng = np.random.default_rng(4242)
    rd = GpuReadData()
    for i in range(10):
        t = random_fastq(rng, int(rng.integers(1, 3000)), int(rng.integers(5, 4000)), crlf=(i % 4 == 3),
                         end=("\n", "", "\n@tail", "\n@tail\n", "\n@t\nACGT")[i % 5], short_frac=0.3)
        b, off = orc.fastq_reads(t)
        rd.loadFromText(t)
        check_loaded(rd, orc.store_roundtrip(b), off, f"random text {i}")
    # short-read data: chunks of the pack kernel span hundreds of reads, runs of empty reads
    recs = []
    for i in range(200_000):
        L = int(rng.integers(0, 9)) if i % 7 else 0
        recs.append(b"@\n" + b"ACGT"[i % 4:i % 4 + 1] * L + b"\n+\n" + b"I" * L + b"\n")
    t = b"".join(recs)
    b, off = orc.fastq_reads(t)
    rd.loadFromText(t)
    check_loaded(rd, orc.store_roundtrip(b), off, "200k tiny reads")
    rd.close()


def test_ci_file_from_gz_to_golden_checksums(orc, c1_reads, c1_golden, tmp_path):
    """util/test_file.fastq.gz rebuilt from the committed reads (headers as in the original, qualities
    made up), gzip'ed, loaded by the device loader and pushed through initialize + both bulk queries:
    the checksums are the ones the unmodified reference produced."""
    bases, offsets = c1_reads
    pz = tmp_path / "test_file.fastq.gz"
    with gzip.open(pz, "wb", compresslevel=1) as f:
        f.write(fastq_text_of(bases, offsets))
    rd = GpuReadData()
    rd.loadFromFile(str(pz), GpuReadData.GZIP, low_mem=True)
    check_loaded(rd, bases, offsets, "C1")
    for cfg in c1_golden["settings"][:2]:
        k, n, thr = cfg["k"], cfg["n"], cfg["thr"]
        f = MinHashReadFilter()
        f.k, f.n, f.overlapSketchThreshold = k, n, thr
        f.randNumbers = ns.rand_from_seed(cfg["seed"], n)
        f.initialize(rd)                      # adopts the device-resident reads, no second copy
        assert f"{orc.fnv_u64(f.sketches().ravel()):016x}" == cfg["fnv_sketches"]
        off, ids = f.queryAll(False)
        assert int(off[-1]) == cfg["fwd_total"] and f"{orc.fnv_csr(off, ids):016x}" == cfg["fwd_fnv"]
        off, ids = f.queryAll(True)
        assert int(off[-1]) == cfg["rc_total"] and f"{orc.fnv_csr(off, ids):016x}" == cfg["rc_fnv"]
        # an online query through the same handle
        q = rd.getRead(4)
        want = f.getFilteredReads(q)
        assert 4 in want.tolist()
        f.close()
    assert rd.getRead(4) == bases[int(offsets[4]):int(offsets[5])].tobytes()     # rd survives the filters
    rd.close()


def test_same_stream_as_ascii_load_full_size_reads():
    """20 000 synthetic ~10 kb reads (0.2 Gbases, 0.4 GB of FASTQ text): the device loader and
    nsmh_load_reads_ascii must leave the same reads behind - same offsets, same unpacked bases, same
    sketches and candidate sets."""
    lengths = ns.synth_lengths(20_000, 10_000, seed=5)
    lengths[:3] = [0, 1, 22]
    host = ns.synth_reads_host(lengths, ns.synth_params(genome_len=20_000_000))
    # FASTQ text built with numpy: "@\n" + seq + "\n+\n" + qual + "\n"
    L = lengths.astype(np.int64)
    rec = 2 + L + 3 + L + 1
    start = np.concatenate([[0], np.cumsum(rec)])
    text = np.full(int(start[-1]), ord("I"), dtype=np.uint8)
    text[start[:-1]] = ord("@")
    text[start[:-1] + 1] = 10
    seq0 = start[:-1] + 2
    idx = np.repeat(seq0 - host.offsets[:-1].astype(np.int64), L) + np.arange(int(host.offsets[-1]), dtype=np.int64)
    text[idx] = host.bases
    text[seq0 + L] = 10
    text[seq0 + L + 1] = ord("+")
    text[seq0 + L + 2] = 10
    text[start[1:] - 1] = 10
    rd = GpuReadData()
    rd.loadFromText(text)
    assert (rd.offsets == host.offsets).all()
    assert (rd.getReads(0, rd.getNumReads()) == host.bases).all()
    st = rd.stats()
    assert st["fastq_parse_ms"] > 0 and st["fastq_pack_ms"] > 0
    rnd = ns.rand_from_seed(20261017, 60)
    a, b = MinHashReadFilter(), MinHashReadFilter()
    for f in (a, b):
        f.k, f.n, f.overlapSketchThreshold, f.randNumbers = 23, 60, 6, rnd
    a.initialize(rd)
    b.initialize(host)
    assert (a.sketches() == b.sketches()).all()
    oa, ia = a.queryAll(False)
    ob, ib = b.queryAll(False)
    assert (oa == ob).all() and (ia == ib).all()
    a.close()
    b.close()
    rd.close()


def test_set_params_keeps_reads(edge):
    """nsmh_set_params: the reads stay on the device, sketches / tables follow the new parameters."""
    host = ReadData(edge["bases"], edge["offsets"])
    rd = GpuReadData()
    rd.loadFromText(fastq_text_of(host.bases, host.offsets))
    for ci in (0, 2, 1):
        seed, k, n, thr = (int(x) for x in edge["cfgs"][ci])
        f = MinHashReadFilter()
        f.k, f.n, f.overlapSketchThreshold, f.randNumbers = k, n, thr, edge[f"rand_{ci}"]
        f.initialize(rd)
        assert (f.sketches() == edge[f"sketches_{ci}"]).all(), f"cfg {ci}"
        off, ids = f.queryAll(False)
        assert (off == edge[f"fwd_off_{ci}"]).all() and (ids == edge[f"fwd_ids_{ci}"]).all()
        f.close()
    rd.close()
