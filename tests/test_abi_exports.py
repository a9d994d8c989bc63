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


def test_cpp_mirror_of_the_kmer_counter_seam_compiles_and_links(tmp_path):
    """include/btgpu.hpp mirrors the reference's KmerCounter seam (KmerCounter.hpp:53-67: findVariantClusterPaths, countPathKmers,
    countInterclusterKmers, parseSampleKmers, classifyPathKmers) over btg_graphs / btg_counter: every method instantiated, compiled with
    -Wall -Wextra and linked against the library (nothing is run: no GPU here)."""
    import subprocess
    from pathlib import Path
    build.build_lib()
    root = Path(__file__).resolve().parent.parent
    exe = tmp_path / "mirror_check"
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I", str(root / "include"), str(root / "tests" / "host_counter_mirror_check.cpp"),
                           "-L", str(capi.LIB_PATH.parent), "-lbtgpu", f"-Wl,-rpath,{capi.LIB_PATH.parent}", "-o", str(exe)])
    assert subprocess.run([str(exe)]).returncode == 0
