"""ctypes binding of libnsmh.so (include/nsmh.h).  Fails loudly when the CUDA
library has not been built: there is no CPU fallback in this package."""
import ctypes as C
import os
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIB_PATH = os.path.join(PKG, "libnsmh.so")
HEADER = os.path.join(ROOT, "include", "nsmh.h")

MG_TOKEN_BYTES, MG_MAX_RANKS = 256, 16
NSMH_OK, NSMH_EINVAL, NSMH_ECUDA, NSMH_ESTATE, NSMH_ENOMEM, NSMH_ERANGE = 0, -1, -2, -3, -4, -5
FLAG_REPETITIVE, FLAG_SHORT = 1, 2

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)


class NsmhError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"nsmh error {code}: {msg}")
        self.code = code


class Stats(C.Structure):
    _fields_ = [("h2d_pack_ms", C.c_float), ("pack_ms", C.c_float), ("sketch_ms", C.c_float),
                ("sketch_main_ms", C.c_float), ("build_ms", C.c_float), ("query_ms", C.c_float),
                ("sketch_fixups", C.c_uint64), ("query_pairs", C.c_uint64),
                ("kernel_launches", C.c_uint32), ("fastq_parse_ms", C.c_float), ("fastq_pack_ms", C.c_float),
                ("fastq_load_ms", C.c_float), ("query_heavy", C.c_uint32), ("query_sorted", C.c_uint32)]


class SynthParams(C.Structure):
    _fields_ = [("genome_len", C.c_uint64), ("genome_seed", C.c_uint64), ("read_seed", C.c_uint64),
                ("p_ins", C.c_float), ("p_del", C.c_float), ("p_sub", C.c_float), ("p_rc", C.c_float)]


def build(force=False):
    """Compile libnsmh.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    src = os.path.join(PKG, "csrc")
    if force:
        subprocess.check_call(["make", "-s", "-C", src, "clean"])
    subprocess.check_call(["make", "-s", "-j8", "-C", src, "all"])
    return LIB_PATH


_SIGS = {
    "nsmh_create": [C.c_uint32, C.c_uint32, C.c_uint32, u64p, C.c_int, C.POINTER(C.c_void_p)],
    "nsmh_destroy": [C.c_void_p],
    "nsmh_rand_from_seed": [C.c_uint32, C.c_uint32, u64p],
    "nsmh_host_alloc": [C.c_size_t, C.POINTER(C.c_void_p)],
    "nsmh_host_free": [C.c_void_p],
    "nsmh_host_alloc_near": [C.c_int, C.c_size_t, C.POINTER(C.c_void_p)],
    "nsmh_bind_thread_near": [C.c_int, C.POINTER(C.c_int)],
    "nsmh_load_reads_ascii": [C.c_void_p, C.c_void_p, u64p, C.c_uint32],
    "nsmh_load_reads_ascii_device": [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64],
    "nsmh_load_reads_dnabitset": [C.c_void_p, C.c_void_p, u32p, C.c_uint32],
    "nsmh_load_sketch_ascii": [C.c_void_p, C.c_void_p, u64p, C.c_uint32],
    "nsmh_load_sketch_dnabitset": [C.c_void_p, C.c_void_p, u32p, C.c_uint32],
    "nsmh_initialize_ascii": [C.c_void_p, C.c_void_p, u64p, C.c_uint32],
    "nsmh_initialize_dnabitset": [C.c_void_p, C.c_void_p, u32p, C.c_uint32],
    "nsmh_num_reads": [C.c_void_p, u32p, u64p],
    "nsmh_set_params": [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, u64p],
    "nsmh_load_fastq": [C.c_void_p, C.c_void_p, C.c_size_t],
    "nsmh_load_fastq_device": [C.c_void_p, C.c_void_p, C.c_size_t],
    "nsmh_load_fastq_file": [C.c_void_p, C.c_char_p, C.c_int],
    "nsmh_read_offsets": [C.c_void_p, u64p],
    "nsmh_get_reads_ascii": [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p],
    "nsmh_read_flags": [C.c_void_p, u8p],
    "nsmh_read_flags_device_ptr": [C.c_void_p, C.POINTER(C.c_void_p)],
    "nsmh_query_all_drop": [C.c_void_p, C.c_uint32, u64p],
    "nsmh_sketch": [C.c_void_p],
    "nsmh_set_sketch_mode": [C.c_void_p, C.c_int],
    "nsmh_get_sketches": [C.c_void_p, u64p],
    "nsmh_sketches_device_ptr": [C.c_void_p, C.POINTER(C.c_void_p)],
    "nsmh_set_table_sketches": [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32],
    "nsmh_build": [C.c_void_p],
    "nsmh_sketch_build": [C.c_void_p],
    "nsmh_table_num_keys": [C.c_void_p, C.c_uint32, u32p],
    "nsmh_query_all": [C.c_void_p, C.c_int, u64p],
    "nsmh_query_all_result": [C.c_void_p, u64p, u32p],
    "nsmh_query_all_device_ptrs": [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)],
    "nsmh_probe_lists": [C.c_void_p, C.c_void_p, C.c_uint32, u64p],
    "nsmh_count_lists": [C.c_void_p, C.c_uint32, C.c_uint32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), u64p],
    "nsmh_mg_init": [C.c_void_p, C.c_uint32, C.c_uint32, u32p, C.c_void_p],
    "nsmh_mg_connect": [C.c_void_p, C.c_void_p],
    "nsmh_mg_run": [C.c_void_p, u64p],
    "nsmh_mg_sketch_run": [C.c_void_p, u64p],
    "nsmh_mg_stage_ms": [C.c_void_p, C.POINTER(C.c_float)],
    "nsmh_mg_shutdown": [C.c_void_p],
    "nsmh_multi_create": [C.c_uint32, C.c_uint32, C.c_uint32, u64p, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)],
    "nsmh_multi_destroy": [C.c_void_p],
    "nsmh_multi_num_devices": [C.c_void_p, C.POINTER(C.c_int)],
    "nsmh_multi_load_reads_ascii": [C.c_void_p, C.c_void_p, u64p, C.c_uint32],
    "nsmh_multi_load_reads_dnabitset": [C.c_void_p, C.c_void_p, u32p, C.c_uint32],
    "nsmh_multi_num_reads": [C.c_void_p, u32p, u64p],
    "nsmh_multi_shards": [C.c_void_p, u32p],
    "nsmh_multi_sketch": [C.c_void_p],
    "nsmh_multi_get_sketches": [C.c_void_p, u64p],
    "nsmh_multi_build": [C.c_void_p],
    "nsmh_multi_query_string": [C.c_void_p, C.c_char_p, C.c_size_t, u32p, C.c_size_t, C.POINTER(C.c_size_t)],
    "nsmh_multi_query_all": [C.c_void_p, C.c_int, u64p],
    "nsmh_multi_query_all_result": [C.c_void_p, u64p, u32p],
    "nsmh_query_string": [C.c_void_p, C.c_char_p, C.c_size_t, u32p, C.c_size_t, C.POINTER(C.c_size_t)],
    "nsmh_query_strings": [C.c_void_p, C.c_void_p, u64p, C.c_uint32, u64p, u32p, C.c_size_t],
    "nsmh_query_sketches": [C.c_void_p, u64p, C.c_uint32, u64p, u32p, C.c_size_t],
    "nsmh_get_stats": [C.c_void_p, C.POINTER(Stats)],
    "nsmh_stream": [C.c_void_p, C.POINTER(C.c_void_p)],
    "nsmh_synchronize": [C.c_void_p],
    "nsmh_synth_reads_host": [C.POINTER(SynthParams), C.c_uint64, C.c_uint32, u64p, C.c_void_p],
    "nsmh_synth_reads_device": [C.c_int, C.POINTER(SynthParams), C.c_uint64, C.c_uint32, C.c_void_p,
                                C.c_void_p],
}

_lib = None


def lib():
    """The loaded library; raises if it is missing (no silent fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a). nanospring_b200 has no CPU or PyTorch fallback.")
        L = C.CDLL(LIB_PATH)
        for name, args in _SIGS.items():
            fn = getattr(L, name)
            fn.argtypes = args
            fn.restype = C.c_int
        L.nsmh_last_error.restype = C.c_char_p
        L.nsmh_last_error.argtypes = []
        L.nsmh_version.restype = C.c_char_p
        L.nsmh_version.argtypes = []
        _lib = L
    return _lib


def check(rc):
    if rc != NSMH_OK:
        raise NsmhError(rc, lib().nsmh_last_error().decode(errors="replace"))
    return rc


def exported_symbols():
    """Names declared in include/nsmh.h (every `int nsmh_*(` / `const char *nsmh_*(`)."""
    import re
    text = open(HEADER).read()
    return sorted(set(re.findall(r"\b(nsmh_[a-z0-9_]+)\s*\(", text)))
