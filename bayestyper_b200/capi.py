"""ctypes binding of libbtgpu.so — the same C ABI (include/btgpu.h) a BayesTyper
maintainer would bind from C++.  There is NO CPU fallback: if the shared library
is missing or no B200 is visible, loading / btg_init fails loudly."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
LIB_PATH = Path(os.environ["BTG_LIB"]) if os.environ.get("BTG_LIB") else ROOT / "bayestyper_b200" / "lib" / "libbtgpu.so"   # BTG_LIB: build variants for experiments

u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
vp = C.c_void_p


class BtgError(RuntimeError):
    pass


_lib = None


def load(build_if_missing: bool = False) -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if build_if_missing:
            from . import build
            build.build_lib()
        else:
            raise BtgError(f"{LIB_PATH} not built (run `python -m bayestyper_b200.build`); there is no CPU fallback")
    L = C.CDLL(str(LIB_PATH))
    L.btg_last_error.restype = C.c_char_p
    L.btg_init.argtypes = [C.c_int]
    L.btg_host_alloc.restype = vp
    L.btg_host_alloc.argtypes = [C.c_size_t]
    L.btg_host_free.argtypes = [vp]
    L.btg_launch_count.restype = C.c_uint64
    # bloom
    L.btg_bloom_create.restype = vp
    L.btg_bloom_create.argtypes = [C.c_uint64, C.c_float, C.c_int]
    L.btg_bloom_load.restype = vp
    L.btg_bloom_load.argtypes = [C.c_char_p, C.c_int]
    L.btg_bloom_from_bytes.restype = vp
    L.btg_bloom_from_bytes.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_int]
    L.btg_bloom_save.argtypes = [vp, C.c_char_p]
    L.btg_bloom_info.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
    L.btg_bloom_download.argtypes = [vp, vp, C.c_uint64]
    L.btg_bloom_insert.argtypes = [vp, vp, C.c_size_t]
    L.btg_bloom_lookup.argtypes = [vp, vp, C.c_size_t, vp]
    L.btg_bloom_insert_dev.argtypes = [vp, vp, C.c_size_t, vp]
    L.btg_bloom_lookup_dev.argtypes = [vp, vp, C.c_size_t, vp, vp]
    L.btg_bloom_lookup_probes_dev.argtypes = [vp, vp, C.c_size_t, vp, vp, vp]
    L.btg_bloom_free.argtypes = [vp]
    L.btg_tbloom_create.restype = vp
    L.btg_tbloom_create.argtypes = [C.c_uint64, C.c_float, C.c_int]
    L.btg_tbloom_info.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
    L.btg_tbloom_insert.argtypes = [vp, vp, C.c_size_t]
    L.btg_tbloom_lookup.argtypes = [vp, vp, C.c_size_t, vp]
    L.btg_tbloom_insert_dev.argtypes = [vp, vp, C.c_size_t, vp]
    L.btg_tbloom_lookup_dev.argtypes = [vp, vp, C.c_size_t, vp, vp]
    L.btg_tbloom_download.argtypes = [vp, vp, C.c_uint64]
    L.btg_tbloom_free.argtypes = [vp]
    L.btg_kmer_hash.argtypes = [vp, C.c_size_t, vp]
    L.btg_kmer_canonical.argtypes = [vp, C.c_size_t, vp]
    L.btg_scan_sequence.argtypes = [C.c_char_p, C.c_size_t, vp, vp]
    L.btg_scan_sequence_dev.argtypes = [vp, C.c_size_t, vp, vp, vp]
    L.btg_scan_sequence_lookup_dev.argtypes = [vp, vp, C.c_size_t, vp, vp]
    _bind_optional(L)
    _lib = L
    return L


def _bind_optional(L):
    """Entry points added by later build stages (bound when present)."""
    from . import capi_ext
    capi_ext.bind(L)


def check(rc, lib=None):
    if rc is None or (isinstance(rc, int) and rc < 0):
        lib = lib or load()
        raise BtgError(lib.btg_last_error().decode())
    return rc


def ptr(a: np.ndarray) -> int:
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def declared_symbols() -> list[str]:
    """Every function name declared in include/btgpu.h."""
    import re
    txt = (ROOT / "include" / "btgpu.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(btg_[a-z0-9_]+)\s*\(", txt)))
