"""In-tree build of libbtgpu.so (nvcc, sm_100a only) and of the oracle checkers.

`python -m bayestyper_b200.build` or `__graft_entry__.build()`.
The .so files are git-ignored but travel to the GPU box with the snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "bayestyper_b200" / "csrc"
LIBDIR = ROOT / "bayestyper_b200" / "lib"
LIB = LIBDIR / "libbtgpu.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
]


# The Gibbs path restates the reference's f32/f64 expressions; the reference is built without FMA contraction
# (x86-64 baseline), so the sampler is too: a contracted a*b+c rounds once instead of twice and shifts e.g. the
# float-typed Gamma parameters of CountDistribution.cpp:182 by one ulp.
PER_FILE_FLAGS = {"gibbs.cu": ["-fmad=false"], "gibbs_wide.cu": ["-fmad=false"]}


def _newer(target: Path, sources) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(Path(s).stat().st_mtime <= t for s in sources)


def build_lib(force: bool = False, verbose: bool = False) -> Path:
    LIBDIR.mkdir(parents=True, exist_ok=True)
    cus = sorted(CSRC.glob("*.cu"))
    deps = cus + sorted(CSRC.glob("*.cuh")) + sorted(CSRC.glob("*.h")) + [ROOT / "include" / "btgpu.h"]
    if not force and _newer(LIB, deps):
        return LIB
    if shutil.which(NVCC) is None and not Path(NVCC).exists():
        raise RuntimeError("nvcc not found; libbtgpu.so cannot be built (no CPU fallback exists)")
    objdir = ROOT / "build" / "obj"
    objdir.mkdir(parents=True, exist_ok=True)
    objs = []
    procs = []
    for cu in cus:
        obj = objdir / (cu.stem + ".o")
        objs.append(obj)
        if not force and _newer(obj, [cu] + [d for d in deps if d.suffix in (".cuh", ".h")]):
            continue
        cmd = [NVCC, *NVCC_FLAGS, *PER_FILE_FLAGS.get(cu.name, []), *os.environ.get("BTG_EXTRA_NVCC_FLAGS", "").split(), "-I", str(ROOT / "include"), "-c", str(cu), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cu, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cu, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {cu}")
    cmd = [NVCC, "-shared", "-o", str(LIB), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fPIC"]
    subprocess.check_call(cmd)
    return LIB


def build_host(force: bool = False) -> Path:
    """g++ build of the C++ host programs over the C ABI: host/btgenotype (Gibbs stage order, links libbtgpu.so), host/btpipeline (both hot paths end to end) and host/btvcf
    (GenotypeWriter, host-only), host/btkmc (KMC database listing, makeBloom), host/btcluster (cluster / group / graph construction, host-only)."""
    hdrs = [ROOT / "include" / "btgpu.hpp", ROOT / "include" / "btgpu.h", ROOT / "include" / "btgpu_vcf.hpp", ROOT / "include" / "btgpu_params.hpp", ROOT / "host" / "btd.hpp", ROOT / "host" / "vcf_desc.hpp"]
    inc = ["-I", str(ROOT / "include"), "-I", str(ROOT / "host")]
    exe = ROOT / "host" / "btgenotype"
    if force or not _newer(exe, [ROOT / "host" / "btgenotype.cpp", LIB, *hdrs]):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", *inc, str(ROOT / "host" / "btgenotype.cpp"), "-L", str(LIBDIR), "-lbtgpu",
                               "-Wl,-rpath,$ORIGIN/../bayestyper_b200/lib", "-o", str(exe)])
    kmc = ROOT / "host" / "btkmc"
    if force or not _newer(kmc, [ROOT / "host" / "btkmc.cpp", ROOT / "include" / "btgpu_kmc.hpp", LIB, *hdrs]):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", *inc, str(ROOT / "host" / "btkmc.cpp"), "-L", str(LIBDIR), "-lbtgpu",
                               "-Wl,-rpath,$ORIGIN/../bayestyper_b200/lib", "-o", str(kmc)])
    pipe = ROOT / "host" / "btpipeline"
    if force or not _newer(pipe, [ROOT / "host" / "btpipeline.cpp", LIB, *hdrs]):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", *inc, str(ROOT / "host" / "btpipeline.cpp"), "-L", str(LIBDIR), "-lbtgpu",
                               "-Wl,-rpath,$ORIGIN/../bayestyper_b200/lib", "-o", str(pipe)])
    vcf = ROOT / "host" / "btvcf"
    if force or not _newer(vcf, [ROOT / "host" / "btvcf.cpp", *hdrs]):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", *inc, str(ROOT / "host" / "btvcf.cpp"), "-o", str(vcf)])
    clu = ROOT / "host" / "btcluster"
    if force or not _newer(clu, [ROOT / "host" / "btcluster.cpp", ROOT / "include" / "btgpu_cluster.hpp", *hdrs]):
        try:        # host-only and independent of the library: a box without zlib's header must not take the other tools down with it
            subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", *inc, str(ROOT / "host" / "btcluster.cpp"), "-lz", "-o", str(clu)])
        except subprocess.CalledProcessError as e:
            print(f"build_host: btcluster not built ({e}); graph_builder.build_genome_graphs_native will refuse to run", file=sys.stderr)
    return exe


def build_oracle(force: bool = False) -> Path:
    """gcc build of the CPU restatement (test infrastructure; see oracle/README.md)."""
    odir = ROOT / "oracle"
    subprocess.check_call(["make", "-s", "-C", str(odir)] + (["-B"] if force else []))
    return odir / "libbtoracle.so"


if __name__ == "__main__":
    v = "-v" in sys.argv
    print(build_lib(force="-f" in sys.argv, verbose=v))
    print(build_host())
    print(build_oracle())
