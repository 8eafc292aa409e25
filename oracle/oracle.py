"""TEST INFRASTRUCTURE ONLY: ctypes bindings for the CPU checkers.

* ``Oracle``  -> oracle/liboracle.so  (plain-C restatement, minhash_oracle.c)
* ``RefLib``  -> oracle/_ref/libnsref.so (the reference's own ReadFilter.cpp /
  BBHashMap.cpp / dnaToBits.cpp compiled unmodified, ref/ref_harness.cpp)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  Nothing in nanospring_b200/ does.
"""
import ctypes as C
import gzip
import os
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libnsref.so")
REF_READDATA_SO = os.path.join(HERE, "_ref", "libnsref_readdata.so")

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)


def build(force=False):
    """Compile liboracle.so (and _ref/libnsref.so when /root/reference exists)."""
    if force or not os.path.exists(ORACLE_SO) or (
            os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(HERE, "minhash_oracle.c"))):
        subprocess.check_call(["make", "-s", "-C", HERE, "all"])
    elif os.path.isdir("/root/reference") and not os.path.exists(REF_SO):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def _p(a, t):
    return a.ctypes.data_as(t)


def reads_to_buffers(reads):
    """list of bytes/str -> (bases uint8[total], offsets uint64[N+1])."""
    bs = [r.encode() if isinstance(r, str) else bytes(r) for r in reads]
    offsets = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        offsets[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    bases = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if bs else np.zeros(0, np.uint8)
    return bases, offsets


def load_fastq(path):
    """Second line of every 4-line record (what ReadData.cpp:113-126 keeps)."""
    op = gzip.open if path.endswith(".gz") else open
    reads = []
    with op(path, "rb") as f:
        while True:
            h = f.readline()
            if not h:
                break
            reads.append(f.readline().rstrip(b"\n"))
            f.readline()
            f.readline()
    return reads


def fnv_sketches(sk):
    return Oracle.get().fnv_u64(np.ascontiguousarray(sk, dtype=np.uint64).ravel())


class Oracle:
    _inst = None

    @classmethod
    def get(cls):
        if cls._inst is None:
            cls._inst = cls()
        return cls._inst

    def __init__(self):
        build()
        L = self.lib = C.CDLL(ORACLE_SO)
        L.orc_rand_from_seed.argtypes = [C.c_uint32, C.c_uint32, u64p]
        L.orc_kmer_to_int.restype = C.c_uint64
        L.orc_kmer_to_int.argtypes = [C.c_char_p, C.c_size_t]
        L.orc_string2kmers.restype = C.c_size_t
        L.orc_string2kmers.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32, u64p]
        L.orc_string2sketch.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32, C.c_uint32, u64p, u64p]
        L.orc_sketch_all.argtypes = [C.c_void_p, u64p, C.c_uint32, C.c_uint32, C.c_uint32, u64p, u64p]
        L.orc_reverse_complement.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
        L.orc_build_tables.restype = C.c_void_p
        L.orc_build_tables.argtypes = [u64p, C.c_uint32, C.c_uint32]
        L.orc_free_tables.argtypes = [C.c_void_p]
        L.orc_table_num_keys.restype = C.c_uint32
        L.orc_table_num_keys.argtypes = [C.c_void_p, C.c_uint32]
        L.orc_query_sketch.restype = C.c_void_p
        L.orc_query_sketch.argtypes = [C.c_void_p, u64p, C.c_uint32, C.POINTER(C.c_size_t)]
        L.orc_query_string.restype = C.c_void_p
        L.orc_query_string.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_uint32, u64p,
                                       C.c_uint32, C.POINTER(C.c_size_t)]
        L.orc_query_all.argtypes = [C.c_void_p, C.c_void_p, u64p, u64p, C.c_uint32, u64p, C.c_uint32,
                                    C.c_int, u64p, C.POINTER(C.c_void_p)]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_read_flags.argtypes = [C.c_void_p, u64p, C.c_uint32, C.c_void_p]
        L.orc_fnv1a64_u64.restype = C.c_uint64
        L.orc_fnv1a64_u64.argtypes = [u64p, C.c_size_t, C.c_uint64]
        L.orc_fnv1a64_csr.restype = C.c_uint64
        L.orc_fnv1a64_csr.argtypes = [u64p, u32p, C.c_uint32]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]
        L.orc_fastq_index.restype = C.c_uint64
        L.orc_fastq_index.argtypes = [C.c_void_p, C.c_size_t, u64p, u64p, C.c_uint64]
        L.orc_store_roundtrip.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]

    # -- small helpers -----------------------------------------------------
    def rand_from_seed(self, seed, n):
        out = np.zeros(n, dtype=np.uint64)
        self.lib.orc_rand_from_seed(seed, n, _p(out, u64p))
        return out

    def kmer_to_int(self, s):
        s = s.encode() if isinstance(s, str) else s
        return int(self.lib.orc_kmer_to_int(s, len(s)))

    def string2kmers(self, s, k):
        s = s.encode() if isinstance(s, str) else s
        out = np.zeros(max(len(s), 1), dtype=np.uint64)
        c = self.lib.orc_string2kmers(s, len(s), k, _p(out, u64p))
        return out[:c].copy()

    def string2sketch(self, s, k, n, rnd):
        s = s.encode() if isinstance(s, str) else s
        sk = np.zeros(n, dtype=np.uint64)
        rnd = np.ascontiguousarray(rnd, dtype=np.uint64)
        self.lib.orc_string2sketch(s, len(s), k, n, _p(rnd, u64p), _p(sk, u64p))
        return sk

    def reverse_complement(self, s):
        s = s.encode() if isinstance(s, str) else bytes(s)
        out = C.create_string_buffer(len(s) + 1)
        self.lib.orc_reverse_complement(s, len(s), out)
        return out.raw[:len(s)]

    def fnv_u64(self, a, h=0xcbf29ce484222325):
        a = np.ascontiguousarray(a, dtype=np.uint64)
        return int(self.lib.orc_fnv1a64_u64(_p(a, u64p), a.size, h))

    def fnv_csr(self, offsets, ids):
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        ids = np.ascontiguousarray(ids, dtype=np.uint32)
        if ids.size == 0:
            ids = np.zeros(1, np.uint32)
        return int(self.lib.orc_fnv1a64_csr(_p(offsets, u64p), _p(ids, u32p), offsets.size - 1))

    def set_num_threads(self, t):
        self.lib.orc_set_num_threads(t)

    def num_threads(self):
        return int(self.lib.orc_num_threads())

    # -- FASTQ ingest ---------------------------------------
    def fastq_index(self, text):
        """(start, len) byte ranges of the reads of a FASTQ text, ReadData.cpp:177-198."""
        t = np.frombuffer(bytes(text), dtype=np.uint8) if not isinstance(text, np.ndarray) else np.ascontiguousarray(text, np.uint8)
        cap = int((t == 10).sum()) // 4 + 2
        start = np.zeros(cap, dtype=np.uint64)
        ln = np.zeros(cap, dtype=np.uint64)
        n = int(self.lib.orc_fastq_index(t.ctypes.data if t.size else None, t.size, _p(start, u64p), _p(ln, u64p), cap))
        assert n <= cap
        return start[:n].copy(), ln[:n].copy()

    def fastq_reads(self, text):
        """(bases uint8[total], offsets uint64[N+1]): the bytes the loader stores, in file order."""
        t = np.frombuffer(bytes(text), dtype=np.uint8) if not isinstance(text, np.ndarray) else np.ascontiguousarray(text, np.uint8)
        start, ln = self.fastq_index(t)
        offsets = np.zeros(start.size + 1, dtype=np.uint64)
        offsets[1:] = np.cumsum(ln, dtype=np.uint64)
        if start.size == 0 or int(offsets[-1]) == 0:
            return np.zeros(0, np.uint8), offsets
        # gather: index of every stored byte = its read's start + position inside the read
        reps = ln.astype(np.int64)
        src = np.repeat(start.astype(np.int64) - offsets[:-1].astype(np.int64), reps) + np.arange(int(offsets[-1]), dtype=np.int64)
        return t[src].copy(), offsets

    def store_roundtrip(self, bases):
        """"ATCG"[code] of every byte: what ReadData::getRead returns (dnaToBits.cpp:46-98)."""
        b = np.ascontiguousarray(bases, dtype=np.uint8)
        out = np.zeros(b.size, dtype=np.uint8)
        if b.size:
            self.lib.orc_store_roundtrip(b.ctypes.data, b.size, out.ctypes.data)
        return out

    # -- bulk --------------------------------------------------------------
    def sketch_all(self, bases, offsets, k, n, rnd):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        rnd = np.ascontiguousarray(rnd, dtype=np.uint64)
        N = offsets.size - 1
        sk = np.zeros((N, n), dtype=np.uint64)
        self.lib.orc_sketch_all(bases.ctypes.data, _p(offsets, u64p), N, k, n, _p(rnd, u64p),
                                _p(sk, u64p))
        return sk

    def read_flags(self, bases, offsets):
        """bit 0: Consensus::checkRepetitive (Consensus.cpp:405-424); bit 1: len < 32 (Consensus.cpp:213)."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        out = np.zeros(offsets.size - 1, dtype=np.uint8)
        self.lib.orc_read_flags(bases.ctypes.data, _p(offsets, u64p), offsets.size - 1, out.ctypes.data)
        return out

    def build_tables(self, sketches):
        sk = np.ascontiguousarray(sketches, dtype=np.uint64)
        N, n = sk.shape
        return OracleTables(self, self.lib.orc_build_tables(_p(sk, u64p), N, n), N, n)


class OracleTables:
    def __init__(self, orc, handle, N, n):
        self.orc, self.h, self.N, self.n = orc, handle, N, n

    def __del__(self):
        if getattr(self, "h", None):
            self.orc.lib.orc_free_tables(self.h)
            self.h = None

    def num_keys(self, j):
        return int(self.orc.lib.orc_table_num_keys(self.h, j))

    def _take(self, ptr, cnt):
        out = np.ctypeslib.as_array(C.cast(ptr, u32p), shape=(max(cnt, 1),))[:cnt].copy()
        self.orc.lib.orc_free(ptr)
        return out

    def query_sketch(self, sketch, thr):
        sk = np.ascontiguousarray(sketch, dtype=np.uint64)
        cnt = C.c_size_t(0)
        ptr = self.orc.lib.orc_query_sketch(self.h, _p(sk, u64p), thr, C.byref(cnt))
        return self._take(ptr, cnt.value)

    def query_string(self, s, k, rnd, thr):
        s = s.encode() if isinstance(s, str) else bytes(s)
        rnd = np.ascontiguousarray(rnd, dtype=np.uint64)
        cnt = C.c_size_t(0)
        ptr = self.orc.lib.orc_query_string(self.h, s, len(s), k, _p(rnd, u64p), thr, C.byref(cnt))
        return self._take(ptr, cnt.value)

    def query_all(self, bases, offsets, sketches, k, rnd, thr, rc):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        sk = np.ascontiguousarray(sketches, dtype=np.uint64)
        rnd = np.ascontiguousarray(rnd, dtype=np.uint64)
        out_off = np.zeros(self.N + 1, dtype=np.uint64)
        ids_ptr = C.c_void_p()
        self.orc.lib.orc_query_all(self.h, bases.ctypes.data, _p(offsets, u64p), _p(sk, u64p), k,
                                   _p(rnd, u64p), thr, int(rc), _p(out_off, u64p), C.byref(ids_ptr))
        return out_off, self._take(ids_ptr.value, int(out_off[-1]))


REF_CONS_SO = os.path.join(HERE, "_ref", "libnsref_consensus.so")


class RefConsensus:
    """The reference's own consensus translation unit (oracle/_ref/libnsref_consensus.so):
    Consensus::initialize / checkRepetitive (Consensus.cpp:405-442)."""
    _inst = None

    @staticmethod
    def available():
        return os.path.exists(REF_CONS_SO)

    @classmethod
    def get(cls):
        if cls._inst is None:
            cls._inst = cls()
        return cls._inst

    def __init__(self):
        L = self.lib = C.CDLL(REF_CONS_SO)
        L.nsref_consensus_is_repetitive.argtypes = [C.c_void_p, u64p, C.c_uint32, C.c_int, C.c_void_p]
        L.nsref_check_repetitive.argtypes = [C.c_char_p, C.c_size_t]

    def is_repetitive(self, bases, offsets, threads=0):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        out = np.zeros(offsets.size - 1, dtype=np.uint8)
        rc = self.lib.nsref_consensus_is_repetitive(bases.ctypes.data, _p(offsets, u64p), offsets.size - 1, threads,
                                                    out.ctypes.data)
        if rc != 0:
            raise RuntimeError("reference consensus harness failed")
        return out

    def check_repetitive(self, s):
        s = s.encode() if isinstance(s, str) else bytes(s)
        return bool(self.lib.nsref_check_repetitive(s, len(s)))


class RefReadData:
    """The reference's own read loader (oracle/_ref/libnsref_readdata.so = unmodified
    src/ReadData.cpp + src/dnaToBits.cpp, oracle/ref/readdata_harness.cpp)."""
    _inst = None

    @staticmethod
    def available():
        return os.path.exists(REF_READDATA_SO)

    @classmethod
    def get(cls):
        if cls._inst is None:
            cls._inst = cls()
        return cls._inst

    def __init__(self):
        L = self.lib = C.CDLL(REF_READDATA_SO)
        L.nsrd_load.restype = C.c_void_p
        L.nsrd_load.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_char_p]
        L.nsrd_num_reads.restype = C.c_uint32
        L.nsrd_num_reads.argtypes = [C.c_void_p]
        for f in (L.nsrd_avg_read_len, L.nsrd_max_read_len):
            f.restype = C.c_uint64
            f.argtypes = [C.c_void_p]
        L.nsrd_lengths.argtypes = [C.c_void_p, u64p]
        L.nsrd_all_reads.argtypes = [C.c_void_p, C.c_void_p]
        L.nsrd_free.argtypes = [C.c_void_p]

    def load(self, path, gzip_flag, low_mem=True):
        """ReadData::loadFromFile(path, GZIP|FASTQ, low_mem) then getRead(0..N-1):
        (bases uint8[total], offsets uint64[N+1], avgReadLen, maxReadLen)."""
        with tempfile.TemporaryDirectory() as td:
            h = self.lib.nsrd_load(os.fsencode(path), 2 if gzip_flag else 0, int(bool(low_mem)), os.fsencode(td))
            if not h:
                raise RuntimeError("reference ReadData::loadFromFile failed")
            try:
                n = int(self.lib.nsrd_num_reads(h))
                ln = np.zeros(max(n, 1), dtype=np.uint64)
                self.lib.nsrd_lengths(h, _p(ln, u64p))
                ln = ln[:n]
                offsets = np.zeros(n + 1, dtype=np.uint64)
                offsets[1:] = np.cumsum(ln, dtype=np.uint64)
                bases = np.zeros(int(offsets[-1]) + 1, dtype=np.uint8)
                self.lib.nsrd_all_reads(h, bases.ctypes.data)
                return (bases[:int(offsets[-1])].copy(), offsets, int(self.lib.nsrd_avg_read_len(h)),
                        int(self.lib.nsrd_max_read_len(h)))
            finally:
                self.lib.nsrd_free(h)

    def load_text(self, text, gzip_flag=False, low_mem=True):
        with tempfile.TemporaryDirectory() as td:
            p = os.path.join(td, "in.fastq" + (".gz" if gzip_flag else ""))
            with (gzip.open(p, "wb") if gzip_flag else open(p, "wb")) as f:
                f.write(bytes(text))
            return self.load(p, gzip_flag, low_mem)


class RefLib:
    """The reference's own code (oracle/_ref/libnsref.so)."""
    _inst = None

    @staticmethod
    def available():
        build()
        return os.path.exists(REF_SO)

    @classmethod
    def get(cls):
        if cls._inst is None:
            cls._inst = cls()
        return cls._inst

    def __init__(self):
        build()
        L = self.lib = C.CDLL(REF_SO)
        L.nsref_create.restype = C.c_void_p
        L.nsref_create.argtypes = [C.c_void_p, u64p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                   u64p, C.c_int, C.c_char_p, C.POINTER(C.c_double),
                                   C.POINTER(C.c_double)]
        L.nsref_create_verbatim.restype = C.c_void_p
        L.nsref_create_verbatim.argtypes = [C.c_void_p, u64p, C.c_uint32, C.c_uint32, C.c_uint32,
                                            C.c_uint32, C.c_int, C.c_char_p, u64p]
        L.nsref_get_sketches.argtypes = [C.c_void_p, u64p]
        L.nsref_query_string.restype = C.c_size_t
        L.nsref_query_string.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, u32p, C.c_size_t]
        L.nsref_query_all.argtypes = [C.c_void_p, C.c_int, C.c_int, u64p, C.POINTER(C.c_void_p),
                                      C.POINTER(C.c_double)]
        L.nsref_free.argtypes = [C.c_void_p]
        L.nsref_destroy.argtypes = [C.c_void_p]
        L.nsref_kmer_to_int.restype = C.c_uint64
        L.nsref_kmer_to_int.argtypes = [C.c_char_p, C.c_size_t]
        L.nsref_string2kmers.restype = C.c_size_t
        L.nsref_string2kmers.argtypes = [C.c_char_p, C.c_size_t, C.c_uint32, u64p]
        L.nsref_rand_from_seed.argtypes = [C.c_uint32, C.c_uint32, u64p]

    def rand_from_seed(self, seed, n):
        out = np.zeros(n, dtype=np.uint64)
        self.lib.nsref_rand_from_seed(seed, n, _p(out, u64p))
        return out

    def kmer_to_int(self, s):
        s = s.encode() if isinstance(s, str) else s
        return int(self.lib.nsref_kmer_to_int(s, len(s)))

    def string2kmers(self, s, k):
        s = s.encode() if isinstance(s, str) else s
        out = np.zeros(max(len(s), 1), dtype=np.uint64)
        c = self.lib.nsref_string2kmers(s, len(s), k, _p(out, u64p))
        return out[:c].copy()

    def create(self, bases, offsets, k, n, thr, rnd, threads=0, verbatim=False):
        return RefFilter(self, bases, offsets, k, n, thr, rnd, threads, verbatim)


class RefFilter:
    """One MinHashReadFilter instance of the reference, initialised on the given reads."""

    def __init__(self, ref, bases, offsets, k, n, thr, rnd, threads, verbatim):
        self.ref = ref
        self.bases = np.ascontiguousarray(bases, dtype=np.uint8)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self.N = self.offsets.size - 1
        self.k, self.n, self.thr = k, n, thr
        self.tmp = tempfile.mkdtemp(prefix="nsref_")
        self.sketch_ms = self.build_ms = None
        if verbatim:
            self.rand = np.zeros(n, dtype=np.uint64)
            self.h = ref.lib.nsref_create_verbatim(self.bases.ctypes.data, _p(self.offsets, u64p),
                                                   self.N, k, n, thr, threads, self.tmp.encode(),
                                                   _p(self.rand, u64p))
        else:
            self.rand = np.ascontiguousarray(rnd, dtype=np.uint64)
            a, b = C.c_double(0), C.c_double(0)
            self.h = ref.lib.nsref_create(self.bases.ctypes.data, _p(self.offsets, u64p), self.N, k,
                                          n, thr, _p(self.rand, u64p), threads, self.tmp.encode(),
                                          C.byref(a), C.byref(b))
            self.sketch_ms, self.build_ms = a.value, b.value
        if not self.h:
            raise RuntimeError("reference harness failed to initialise")

    def close(self):
        if getattr(self, "h", None):
            self.ref.lib.nsref_destroy(self.h)
            self.h = None
            try:
                os.rmdir(self.tmp)
            except OSError:
                pass

    __del__ = close

    def sketches(self):
        sk = np.zeros((self.N, self.n), dtype=np.uint64)
        self.ref.lib.nsref_get_sketches(self.h, _p(sk, u64p))
        return sk

    def query_string(self, s):
        s = s.encode() if isinstance(s, str) else bytes(s)
        cap = self.N + 1
        out = np.zeros(cap, dtype=np.uint32)
        c = self.ref.lib.nsref_query_string(self.h, s, len(s), _p(out, u32p), cap)
        return out[:c].copy()

    def query_all(self, mode, threads=0):
        """mode 0: stored sketches (private overload); 1: RC strings; 2: forward strings."""
        off = np.zeros(self.N + 1, dtype=np.uint64)
        ptr = C.c_void_p()
        ms = C.c_double(0)
        rc = self.ref.lib.nsref_query_all(self.h, mode, threads, _p(off, u64p), C.byref(ptr),
                                          C.byref(ms))
        if rc != 0:
            raise RuntimeError("nsref_query_all failed")
        tot = int(off[-1])
        ids = np.ctypeslib.as_array(C.cast(ptr, u32p), shape=(max(tot, 1),))[:tot].copy()
        self.ref.lib.nsref_free(ptr)
        self.query_ms = ms.value
        return off, ids
