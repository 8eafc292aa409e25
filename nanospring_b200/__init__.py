"""nanospring_b200 — B200-native MinHash read-overlap engine, a drop-in for
NanoSpring's ReadFilter / MinHashReadFilter (include/ReadFilter.h).

Layout:
  csrc/      hand-written sm_100a CUDA kernels + the C ABI (include/nsmh.h) -> libnsmh.so
  cpp/       header-only C++ adaptor `GpuMinHashReadFilter : public ReadFilter`
  filter.py  host-side Python mirror of the reference interface (ctypes over the C ABI):
             MinHashReadFilter, ReadData (host buffers), GpuReadData (FASTQ ingest on the device)
  shard.py   read sharding + sketch all-gather across GPUs (torch.distributed plumbing)
"""
from ._lib import NsmhError, build, lib  # noqa: F401
from .filter import (GpuReadData, MinHashReadFilter, MultiGpuMinHashReadFilter, ReadData, rand_from_seed, reverse_complement,  # noqa: F401
                     synth_lengths, synth_params, synth_reads_host)
