"""The C-ABI library loads and exports every symbol include/btgpu.h declares
(no compute calls: this runs without a GPU)."""
import ctypes as C

from bayestyper_b200 import capi, build


def test_library_exports_every_declared_symbol():
    build.build_lib()
    lib = C.CDLL(str(capi.LIB_PATH))
    names = capi.declared_symbols()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_calls_fail_loudly_without_init():
    lib = capi.load(build_if_missing=True)
    import torch
    if torch.cuda.is_available():
        return
    assert lib.btg_init(0) < 0
    assert b"no CPU fallback" in lib.btg_last_error()
    assert lib.btg_bloom_create(10, 0.001, 55) is None
