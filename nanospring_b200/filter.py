"""Host-side mirror of the reference's read-overlap filter interface.

Reference (paths relative to /root/reference):
  class ReadFilter        include/ReadFilter.h:15-30   initialize(ReadData&), getFilteredReads(string, vector<read_t>&)
  class MinHashReadFilter include/ReadFilter.h:33-139  public fields k, n, overlapSketchThreshold, tempDir
  wiring                  src/Compressor.cpp:69-76     fields assigned, then initialize(rD) once
  caller                  src/Consensus.cpp:180-191    getFilteredReads(window) and getFilteredReads(revcomp window)

Same names, same argument meaning, same error behaviour (exceptions).  All compute
runs in libnsmh.so (hand-written sm_100a kernels) through the C ABI of
include/nsmh.h; this module only marshals buffers.  The C++ twin for linking under
the real Consensus class is nanospring_b200/cpp/GpuMinHashReadFilter.h.
"""
import ctypes as C
import os

import numpy as np

from . import _lib
from ._lib import NSMH_ERANGE, NsmhError, SynthParams, Stats, check, lib, u32p, u64p, u8p


def rand_from_seed(seed, n):
    """rand[i] = i-th output of std::mt19937_64(seed): the numbers generateRandomNumbers()
    (ReadFilter.cpp:49-63) draws once std::random_device has returned `seed`."""
    out = np.zeros(n, dtype=np.uint64)
    check(lib().nsmh_rand_from_seed(int(seed) & 0xFFFFFFFF, n, out.ctypes.data_as(u64p)))
    return out


class ReadData:
    """What the filter needs of the reference's ReadData (include/ReadData.h:26-60):
    getNumReads(), getRead(i), avgReadLen, maxReadLen.  Reads are kept as one
    concatenated ASCII buffer + offsets, the layout the C ABI takes."""

    def __init__(self, bases, offsets):
        self.bases = np.ascontiguousarray(bases, dtype=np.uint8)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        if self.offsets.size < 1 or self.offsets[0] != 0:
            raise ValueError("offsets must start at 0")
        lens = np.diff(self.offsets.astype(np.int64))
        self.numReads = self.offsets.size - 1
        self.maxReadLen = int(lens.max()) if lens.size else 0
        self.avgReadLen = int(lens.sum() // max(self.numReads, 1))

    @classmethod
    def from_reads(cls, reads):
        bs = [r.encode() if isinstance(r, str) else bytes(r) for r in reads]
        offsets = np.zeros(len(bs) + 1, dtype=np.uint64)
        if bs:
            offsets[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
        bases = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, np.uint8)
        return cls(bases, offsets)

    def getNumReads(self):
        return self.numReads

    def getRead(self, i):
        return self.bases[int(self.offsets[i]):int(self.offsets[i + 1])].tobytes()


class GpuReadData:
    """The reference's ReadData with the loader on the device.

    Reference: ReadData::loadFromFile(fileName, FASTQ|GZIP, low_mem) (ReadData.cpp:12-26, 156-221)
    parses the FASTQ text on one host thread, 2-bit packs every read (DnaBitset) into a temp file and
    serves getRead() through a mutex + seekg (ReadData.cpp:225-235).  Here the host only reads (and
    inflates) the file; record splitting, the 2-bit pack and getRead's unpack are kernels
    (csrc/fastq.cu) and the packed reads stay on the device, where MinHashReadFilter.initialize(rD)
    sketches them without another copy.  Same public surface: tempDir, avgReadLen, maxReadLen,
    loadFromFile, getNumReads, getRead."""

    FASTQ, READ, GZIP = 0, 1, 2            # enum Filetype, ReadData.h:23

    def __init__(self, device=0):
        self.device = device
        self.tempDir = ""                   # interface parity; no temp file is written
        self.avgReadLen = 0
        self.maxReadLen = 0
        self.numReads = 0
        self.offsets = np.zeros(1, dtype=np.uint64)
        self._h = None

    def _handle(self):
        if self._h is None:
            h = C.c_void_p()
            rnd = np.zeros(1, dtype=np.uint64)      # placeholder parameters; the filter sets the real ones
            check(lib().nsmh_create(23, 1, 1, rnd.ctypes.data_as(u64p), self.device, C.byref(h)))
            self._h = h
        return self._h

    def close(self):
        if self._h is not None:
            lib().nsmh_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _loaded(self):
        n = C.c_uint32(0)
        check(lib().nsmh_num_reads(self._h, C.byref(n), None))
        self.numReads = n.value
        self.offsets = np.zeros(self.numReads + 1, dtype=np.uint64)
        check(lib().nsmh_read_offsets(self._h, self.offsets.ctypes.data_as(u64p)))
        lens = np.diff(self.offsets.astype(np.int64))
        self.maxReadLen = int(lens.max()) if lens.size else 0                 # ReadData.cpp:182-183
        self.avgReadLen = int(lens.sum() // max(self.numReads, 1))             # ReadData.cpp:200

    def loadFromFile(self, fileName, filetype=0, low_mem=False):
        """low_mem is accepted for signature parity: nothing is spilled to disk either way."""
        if filetype not in (self.FASTQ, self.GZIP):
            raise ValueError("GpuReadData: only FASTQ and GZIP inputs (the two the CLI passes, main.cpp:141-145)")
        check(lib().nsmh_load_fastq_file(self._handle(), os.fsencode(fileName), int(filetype == self.GZIP)))
        self._loaded()

    def loadFromText(self, text):
        """The (inflated) FASTQ text as bytes / uint8 array in host memory."""
        t = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray, memoryview)) else \
            np.ascontiguousarray(text, dtype=np.uint8)
        check(lib().nsmh_load_fastq(self._handle(), t.ctypes.data if t.size else None, t.size))
        self._loaded()

    def loadFromDeviceText(self, d_text_ptr, nbytes):
        check(lib().nsmh_load_fastq_device(self._handle(), d_text_ptr, nbytes))
        self._loaded()

    def getNumReads(self):
        return self.numReads

    def getReads(self, first, count):
        """Concatenated getRead(first) .. getRead(first+count-1) as one uint8 array."""
        nb = int(self.offsets[first + count] - self.offsets[first])
        out = np.zeros(max(nb, 1), dtype=np.uint8)
        check(lib().nsmh_get_reads_ascii(self._h, first, count, out.ctypes.data))
        return out[:nb]

    def getRead(self, i):
        """ReadData::getRead: "ATCG"[code] of every stored base (dnaToBits.cpp:81-98)."""
        return self.getReads(i, 1).tobytes()

    def stats(self):
        st = Stats()
        check(lib().nsmh_get_stats(self._h, C.byref(st)))
        return {f: getattr(st, f) for f, _ in Stats._fields_}


class MinHashReadFilter:
    """Drop-in for the reference's MinHashReadFilter, computing on one B200."""

    def __init__(self, device=0):
        self.k = 23                         # main.cpp:57
        self.n = 60                         # main.cpp:59
        self.overlapSketchThreshold = 6     # main.cpp:61
        self.tempDir = ""                   # kept for interface parity; no temp files are used
        self.device = device
        self.randNumbers = None             # injectable; drawn like the reference if left None
        self.sketchMode = 0                 # 0 filtered kernel, 1 brute force (same results)
        self._h = None
        self._rd = None
        self._borrowed = False              # handle owned by a GpuReadData

    # -- lifetime ----------------------------------------------------------------
    def close(self):
        if self._h is not None:
            if not self._borrowed:
                lib().nsmh_destroy(self._h)
            self._h = None
            self._borrowed = False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def generateRandomNumbers(self, n):
        """ReadFilter.cpp:49-63: a fresh 32-bit seed from the OS entropy source."""
        seed = int.from_bytes(os.urandom(4), "little")
        self.randNumbers = rand_from_seed(seed, n)

    def _create(self):
        self.close()
        if self.randNumbers is None:
            self.generateRandomNumbers(self.n)
        rnd = np.ascontiguousarray(self.randNumbers, dtype=np.uint64)
        if rnd.size != self.n:
            raise ValueError("randNumbers must hold n values")
        h = C.c_void_p()
        check(lib().nsmh_create(self.k, self.n, self.overlapSketchThreshold, rnd.ctypes.data_as(u64p),
                                self.device, C.byref(h)))
        self._h = h
        self._made_with = self._params_key()
        check(lib().nsmh_set_sketch_mode(h, int(self.sketchMode)))

    def _params_key(self):
        return (self.k, self.n, self.overlapSketchThreshold, self.device, int(self.sketchMode),
                None if self.randNumbers is None else np.asarray(self.randNumbers, dtype=np.uint64).tobytes())

    def _create_if_changed(self):
        """A handle made with the same k, n, threshold, random numbers and device is kept (its streams, pools
        and filter tables are expensive to remake); anything else gets a fresh one."""
        if self._h is None or getattr(self, "_made_with", None) != self._params_key():
            self._create()

    # -- initialize() split in its stages (bench / multi-GPU drive them separately) ----
    def load(self, rD):
        """Ship the reads to the device and 2-bit pack them there."""
        if not isinstance(rD, ReadData):
            rD = ReadData.from_reads(rD)
        self._rd = rD
        if self._h is None:
            self._create()
        check(lib().nsmh_load_reads_ascii(self._h, rD.bases.ctypes.data, rD.offsets.ctypes.data_as(u64p),
                                          rD.numReads))

    def load_sketch(self, rD=None, packed=None):
        """load() + sketch() as one pipelined call (a chunk is sketched while the next one crosses PCIe), for
        flows that build the tables elsewhere (multi-GPU).  rD: ReadData, or packed=(DnaBitset bytes, u32 lengths)."""
        if self._h is None:
            self._create()
        if packed is not None:
            pk = np.ascontiguousarray(packed[0], dtype=np.uint8)
            ln = np.ascontiguousarray(packed[1], dtype=np.uint32)
            check(lib().nsmh_load_sketch_dnabitset(self._h, pk.ctypes.data, ln.ctypes.data_as(u32p), ln.size))
            return
        if not isinstance(rD, ReadData):
            rD = ReadData.from_reads(rD)
        self._rd = rD
        check(lib().nsmh_load_sketch_ascii(self._h, rD.bases.ctypes.data, rD.offsets.ctypes.data_as(u64p), rD.numReads))

    def load_device(self, d_bases_ptr, d_offsets_ptr, num_reads, total_bases):
        """Reads already resident on the device (ASCII bases + u64 offsets)."""
        if self._h is None:
            self._create()
        check(lib().nsmh_load_reads_ascii_device(self._h, d_bases_ptr, d_offsets_ptr, num_reads,
                                                 total_bases))

    def load_dnabitset(self, packed, lengths):
        """Reads 2-bit packed the reference's way (DnaBitset, dnaToBits.cpp:11-36)."""
        if self._h is None:
            self._create()
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        lengths = np.ascontiguousarray(lengths, dtype=np.uint32)
        check(lib().nsmh_load_reads_dnabitset(self._h, packed.ctypes.data, lengths.ctypes.data_as(u32p),
                                              lengths.size))

    def sketch(self):
        check(lib().nsmh_sketch(self._h))

    def build(self):
        check(lib().nsmh_build(self._h))

    def sketch_build(self):
        """sketch() + build() as one call: the sketch's exact fix-up pass runs beside the table insert."""
        check(lib().nsmh_sketch_build(self._h))

    def _adopt(self, rD):
        """Reads already packed on the device by GpuReadData: take its handle, set k/n/thr/rand."""
        self.close()
        if rD._h is None:
            raise RuntimeError("GpuReadData: nothing loaded")
        if self.randNumbers is None:
            self.generateRandomNumbers(self.n)
        rnd = np.ascontiguousarray(self.randNumbers, dtype=np.uint64)
        if rnd.size != self.n:
            raise ValueError("randNumbers must hold n values")
        check(lib().nsmh_set_params(rD._h, self.k, self.n, self.overlapSketchThreshold, rnd.ctypes.data_as(u64p)))
        self._h, self._borrowed, self._rd = rD._h, True, rD
        check(lib().nsmh_set_sketch_mode(self._h, int(self.sketchMode)))

    def initialize(self, rD):
        """ReadFilter.cpp:11-47: sketch every read, then populate the n hash tables."""
        if isinstance(rD, GpuReadData):
            self._adopt(rD)
            self.sketch()
            self.build()
            return
        if not isinstance(rD, ReadData):
            rD = ReadData.from_reads(rD)
        self._rd = rD
        self._create_if_changed()
        # load + sketch + build as one pipelined call: chunk i is sketched while chunk i+1 crosses PCIe
        check(lib().nsmh_initialize_ascii(self._h, rD.bases.ctypes.data, rD.offsets.ctypes.data_as(u64p),
                                          rD.numReads))

    def initialize_dnabitset(self, packed, lengths):
        """initialize() from the reference's 2-bit store (DnaBitset bytes + u32 lengths, dnaToBits.cpp:11-36):
        what ReadData holds in memory / in its temp file.  Same pipelined call as initialize()."""
        self._create_if_changed()
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        lengths = np.ascontiguousarray(lengths, dtype=np.uint32)
        check(lib().nsmh_initialize_dnabitset(self._h, packed.ctypes.data, lengths.ctypes.data_as(u32p), lengths.size))

    # -- queries -------------------------------------------------------------------
    def getFilteredReads(self, s, results=None):
        """ReadFilter.cpp:85-97.  Returns ascending read ids (np.uint32); if `results` is a
        list it is cleared and filled, like the reference's out-parameter."""
        if self._h is None:
            raise RuntimeError("getFilteredReads before initialize")
        s = s.encode() if isinstance(s, str) else bytes(s)
        cap = 256
        while True:
            out = np.empty(cap, dtype=np.uint32)
            cnt = C.c_size_t(0)
            rc = lib().nsmh_query_string(self._h, s, len(s), out.ctypes.data_as(u32p), cap, C.byref(cnt))
            if rc == NSMH_ERANGE:
                cap = cnt.value
                continue
            check(rc)
            break
        res = out[:cnt.value].copy()
        if results is not None:
            results.clear()
            results.extend(int(x) for x in res)
        return res

    def getFilteredReadsBatch(self, strings):
        """Several strings in one launch sequence (e.g. a window and its reverse complement)."""
        rd = ReadData.from_reads(strings)
        off_out = np.zeros(rd.numReads + 1, dtype=np.uint64)
        cap = 1024
        while True:
            ids = np.empty(cap, dtype=np.uint32)
            rc = lib().nsmh_query_strings(self._h, rd.bases.ctypes.data, rd.offsets.ctypes.data_as(u64p),
                                          rd.numReads, off_out.ctypes.data_as(u64p),
                                          ids.ctypes.data_as(u32p), cap)
            if rc == NSMH_ERANGE:
                cap = int(off_out[-1])
                continue
            check(rc)
            break
        return [ids[int(off_out[i]):int(off_out[i + 1])].copy() for i in range(rd.numReads)]

    def querySketches(self, sketches):
        """getFilteredReads(kMer_t sketch[], results) (ReadFilter.cpp:65-83) for a batch."""
        sk = np.ascontiguousarray(sketches, dtype=np.uint64).reshape(-1, self.n)
        off_out = np.zeros(sk.shape[0] + 1, dtype=np.uint64)
        cap = 1024
        while True:
            ids = np.empty(cap, dtype=np.uint32)
            rc = lib().nsmh_query_sketches(self._h, sk.ctypes.data_as(u64p), sk.shape[0],
                                           off_out.ctypes.data_as(u64p), ids.ctypes.data_as(u32p), cap)
            if rc == NSMH_ERANGE:
                cap = int(off_out[-1])
                continue
            check(rc)
            break
        return off_out, ids[:int(off_out[-1])].copy()

    def queryAll(self, reverseComplement=False, fetch=True):
        """Every loaded read against the tables (bulk form of the loop at Consensus.cpp:185-191).
        Returns (offsets u64[N+1], ids u32[total]) or just the total when fetch=False."""
        total = C.c_uint64(0)
        check(lib().nsmh_query_all(self._h, int(bool(reverseComplement)), C.byref(total)))
        if not fetch:
            return total.value
        N = self.numReads()
        off = np.zeros(N + 1, dtype=np.uint64)
        ids = np.zeros(max(total.value, 1), dtype=np.uint32)
        check(lib().nsmh_query_all_result(self._h, off.ctypes.data_as(u64p), ids.ctypes.data_as(u32p)))
        return off, ids[:total.value]

    # -- candidate pre-filters of the consensus builder (csrc/prefilter.cu) ---------
    def readFlags(self):
        """u8[numReads]: bit 0 = Consensus::checkRepetitive (Consensus.cpp:405-424, i.e. isRepetitive[]
        of Consensus::initialize), bit 1 = shorter than 32 bases (Consensus.cpp:213)."""
        out = np.zeros(max(self.numReads(), 1), dtype=np.uint8)
        check(lib().nsmh_read_flags(self._h, out.ctypes.data_as(u8p)))
        return out[:self.numReads()]

    def queryAllDrop(self, drop_mask=3, fetch=True):
        """Drop the candidates whose flags intersect drop_mask from the CSR of the last queryAll, on
        the device (the `continue`s at Consensus.cpp:204-216)."""
        total = C.c_uint64(0)
        check(lib().nsmh_query_all_drop(self._h, int(drop_mask), C.byref(total)))
        if not fetch:
            return total.value
        N = self.numReads()
        off = np.zeros(N + 1, dtype=np.uint64)
        ids = np.zeros(max(total.value, 1), dtype=np.uint32)
        check(lib().nsmh_query_all_result(self._h, off.ctypes.data_as(u64p), ids.ctypes.data_as(u32p)))
        return off, ids[:total.value]

    # -- accessors -----------------------------------------------------------------
    def numReads(self):
        n = C.c_uint32(0)
        check(lib().nsmh_num_reads(self._h, C.byref(n), None))
        return n.value

    def sketches(self):
        """The [numReads][n] sketch matrix (row-major, like ReadFilter.cpp:21)."""
        N = self.numReads()
        out = np.zeros((N, self.n), dtype=np.uint64)
        check(lib().nsmh_get_sketches(self._h, out.ctypes.data_as(u64p)))
        return out

    def sketchesDevicePtr(self):
        p = C.c_void_p()
        check(lib().nsmh_sketches_device_ptr(self._h, C.byref(p)))
        return p.value

    def setTableSketches(self, d_ptr, table_reads, id_base):
        check(lib().nsmh_set_table_sketches(self._h, d_ptr, table_reads, id_base))

    def tableNumKeys(self, j):
        v = C.c_uint32(0)
        check(lib().nsmh_table_num_keys(self._h, j, C.byref(v)))
        return v.value

    def stats(self):
        st = Stats()
        check(lib().nsmh_get_stats(self._h, C.byref(st)))
        return {f: getattr(st, f) for f, _ in Stats._fields_}

    def stream(self):
        p = C.c_void_p()
        check(lib().nsmh_stream(self._h, C.byref(p)))
        return p.value or 0

    def synchronize(self):
        check(lib().nsmh_synchronize(self._h))


def reverse_complement(s):
    """ReadData::toReverseComplement (ReadData.h:163-172, ReadData.cpp:247-260): only the
    bytes 'A','C','G','T' are complemented."""
    s = s.encode() if isinstance(s, str) else bytes(s)
    return s[::-1].translate(bytes.maketrans(b"ACGT", b"TGCA"))


# ---- synthetic reads (createData.py recipe; see csrc/synth.cu) -------------------------
class MultiGpuMinHashReadFilter:
    """MinHashReadFilter over several GPUs of ONE process (include/nsmh.h, nsmh_multi_*): the reference's
    own process model - initialize() once (Compressor.cpp:69-76), getFilteredReads() from many threads
    (Consensus.cpp:29,189).  Reads are split by bases and sketched in parallel, every device then holds
    the tables of ALL reads and online queries are spread over the devices."""

    def __init__(self, devices=(0,)):
        self.k, self.n, self.overlapSketchThreshold = 23, 60, 6
        self.tempDir = ""
        self.devices = [int(d) for d in devices]
        self.randNumbers = None
        self._m = None

    def close(self):
        if self._m is not None:
            lib().nsmh_multi_destroy(self._m)
            self._m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def initialize(self, rD, packed=None):
        """rD: ReadData (ASCII) - or packed=(DnaBitset bytes, u32 lengths), the reference's 2-bit store."""
        self.close()
        if self.randNumbers is None:
            self.randNumbers = rand_from_seed(int.from_bytes(os.urandom(4), "little"), self.n)
        rnd = np.ascontiguousarray(self.randNumbers, dtype=np.uint64)
        dev = (C.c_int * len(self.devices))(*self.devices)
        m = C.c_void_p()
        check(lib().nsmh_multi_create(self.k, self.n, self.overlapSketchThreshold, rnd.ctypes.data_as(u64p), dev,
                                      len(self.devices), C.byref(m)))
        self._m = m
        if packed is not None:
            pk = np.ascontiguousarray(packed[0], dtype=np.uint8)
            ln = np.ascontiguousarray(packed[1], dtype=np.uint32)
            check(lib().nsmh_multi_load_reads_dnabitset(m, pk.ctypes.data, ln.ctypes.data_as(u32p), ln.size))
        else:
            check(lib().nsmh_multi_load_reads_ascii(m, rD.bases.ctypes.data, rD.offsets.ctypes.data_as(u64p), rD.numReads))
        check(lib().nsmh_multi_sketch(m))
        check(lib().nsmh_multi_build(m))

    def numReads(self):
        nr = C.c_uint32(0)
        check(lib().nsmh_multi_num_reads(self._m, C.byref(nr), None))
        return nr.value

    def shards(self):
        out = np.zeros(len(self.devices) + 1, dtype=np.uint32)
        check(lib().nsmh_multi_shards(self._m, out.ctypes.data_as(u32p)))
        return out

    def sketches(self):
        out = np.zeros((self.numReads(), self.n), dtype=np.uint64)
        check(lib().nsmh_multi_get_sketches(self._m, out.ctypes.data_as(u64p)))
        return out

    def getFilteredReads(self, s, results=None):
        s = s.encode() if isinstance(s, str) else bytes(s)
        cap = 1024
        while True:
            out = np.empty(cap, dtype=np.uint32)
            cnt = C.c_size_t(0)
            rc = lib().nsmh_multi_query_string(self._m, s, len(s), out.ctypes.data_as(u32p), cap, C.byref(cnt))
            if rc == NSMH_ERANGE:
                cap = cnt.value
                continue
            check(rc)
            break
        res = out[:cnt.value].copy()
        if results is not None:
            results.clear()
            results.extend(int(x) for x in res)
        return res

    def queryAll(self, reverseComplement=False):
        total = C.c_uint64(0)
        check(lib().nsmh_multi_query_all(self._m, int(bool(reverseComplement)), C.byref(total)))
        off = np.zeros(self.numReads() + 1, dtype=np.uint64)
        ids = np.zeros(max(total.value, 1), dtype=np.uint32)
        check(lib().nsmh_multi_query_all_result(self._m, off.ctypes.data_as(u64p), ids.ctypes.data_as(u32p)))
        return off, ids[:total.value]


def synth_params(genome_len=50_000_000, genome_seed=1, read_seed=2, p_ins=0.03, p_del=0.03,
                 p_sub=0.04, p_rc=0.5):
    return SynthParams(genome_len, genome_seed, read_seed, p_ins, p_del, p_sub, p_rc)


def synth_lengths(num_reads, mean, seed=2, dist="mixgamma"):
    """Read lengths: the two-gamma mixture of createData.py:42-60 (draw_mix_gamma_dis)
    rescaled to `mean`, clipped to >= 1; `const` gives fixed-length reads."""
    if dist == "const":
        return np.full(num_reads, int(mean), dtype=np.uint64)
    rng = np.random.Generator(np.random.PCG64(seed))
    half = num_reads // 2
    # scipy's gamma.rvs(a, loc) as used there: shape a, location loc, scale 1
    s1 = rng.gamma(6.3693711, 1.0, size=half) + 0.53834893
    s0 = rng.gamma(1.67638771, 1.0, size=num_reads - half) + 0.22871401
    sample = np.concatenate((s0, s1))
    rng.shuffle(sample)
    sample = sample * mean / 4.39          # createData.py:57
    return np.clip(sample.astype(np.int64), 1, None).astype(np.uint64)


def synth_reads_host(lengths, params=None, first_read=0):
    """Synthetic reads on the host: returns ReadData."""
    params = params or synth_params()
    lengths = np.asarray(lengths, dtype=np.uint64)
    offsets = np.zeros(lengths.size + 1, dtype=np.uint64)
    offsets[1:] = np.cumsum(lengths, dtype=np.uint64)
    bases = np.zeros(int(offsets[-1]) + 1, dtype=np.uint8)
    check(lib().nsmh_synth_reads_host(C.byref(params), first_read, lengths.size,
                                      offsets.ctypes.data_as(u64p), bases.ctypes.data))
    return ReadData(bases[:int(offsets[-1])], offsets)
