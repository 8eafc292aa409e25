"""BASELINE.json configs at their FULL sizes on one B200 (tests/config_sweep.py does the work):
configs[1] (100 k reads, ~1 Gbase), configs[2]'s read set (1 M reads, ~10 Gbases) on a single GPU,
the corners of configs[3]'s parameter sweep on it, and configs[4]'s ultra-long reads (~5 Gbases).
No CPU pass over 10 Gbases: parity goes through size-independent properties - brute-force kernel ==
filter kernel bit for bit, sampled sketch rows == CPU oracle, sampled candidate lists == the
reference's definition evaluated independently with torch integer ops on the full sketch matrix,
self-inclusion, strictly ascending lists, symmetry of the forward relation."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _check(res):
    assert res["ok"], res
    assert res["candidate_ids"] >= res["reads"]
    print(res)


def test_c2_100k_reads_low_error_rich_candidates():
    import config_sweep as cs
    r = cs.DeviceReads(100_000, 10_000, 0.02, seed=1000)
    res = cs.run_config("c2lo", r, 23, 60, 6, steps=2, warmup=1)
    _check(res)
    assert res["candidate_ids"] > 5 * res["reads"]          # ~20x coverage at 2 % error: real overlaps


@pytest.fixture(scope="module")
def c3_reads():
    import torch
    import config_sweep as cs
    free, _ = torch.cuda.mem_get_info(0)
    if free < 40e9:
        pytest.skip("needs 40 GB of free device memory")
    return cs.DeviceReads(1_000_000, 10_000, 0.10, seed=2000)


@pytest.mark.parametrize("k,n,thr", [(23, 60, 6), (15, 30, 3), (31, 120, 12)])
def test_c3_c4_one_million_reads(c3_reads, k, n, thr):
    import config_sweep as cs
    assert c3_reads.total > 9.5e9
    _check(cs.run_config(f"c3/c4 k={k} n={n}", c3_reads, k, n, thr, steps=1, warmup=1, sample=32))


def test_c5_ultra_long_reads():
    import config_sweep as cs
    r = cs.DeviceReads(50_000, 100_000, 0.10, seed=3000)
    assert r.total > 4.5e9 and int(r.lengths.max()) > 300_000
    _check(cs.run_config("c5", r, 23, 60, 6, steps=1, warmup=1, sample=16))
