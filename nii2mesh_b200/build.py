"""Build libb2m.so (hand-written sm_100a CUDA kernels + plain-C host side) in-tree with nvcc/gcc.

    python -m nii2mesh_b200.build [--force] [--verbose]

Flags that matter for bit-exactness against the reference: no FMA contraction
(-fmad=false), IEEE division and square root (-prec-div=true -prec-sqrt=true), no flush-to-zero
(-ftz=false), never --use_fast_math.  sm_100a only: no other -gencode, no PTX fallback.
"""
import subprocess
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libb2m.so"
CU = ["ctx.cu", "scan.cu", "smooth.cu", "cc.cu", "mc.cu", "weld.cu", "pipeline.cu", "tables.cu", "comm.cu", "atlas.cu", "isolevel.cu", "io.cu", "post.cu"]
CC = ["meshify_host.c", "hostcopy.c"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-strict-aliasing",
]


def _stale():
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.c")) + list(CSRC.glob("*.inc")) + \
        list((HERE.parent / "include").glob("*.h")) + [Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """defines/out: an experiment build (e.g. defines=["-DSX_Y_F2F=0"], out=HERE / "libb2m_alt.so", loaded through
    the B2M_LIBPATH environment variable) next to the product library"""
    alt = out is not None
    out = Path(out) if alt else LIB
    if not alt and not force and not _stale():
        return LIB
    objdir = HERE / ("build_alt" if alt else "build")
    objdir.mkdir(exist_ok=True)
    objs = []
    procs = []
    for f in CU:
        o = objdir / (f + ".o")
        cmd = ["nvcc", *NVCC_FLAGS, *defines, "-c", str(CSRC / f), "-o", str(o)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((f, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(o))
    for f in CC:
        o = objdir / (f + ".o")
        cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-D_POSIX_C_SOURCE=200809L", "-c", str(CSRC / f), "-o", str(o)]
        procs.append((f, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(o))
    failed = False
    for f, p in procs:
        log, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- {f}\n{log}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("libb2m build failed")
    subprocess.check_call(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(out), *objs,
                           "-lcudart_static", "-lm", "-ldl", "-lrt", "-lpthread", "-Xlinker", "--no-undefined"])
    return out


if __name__ == "__main__":
    defs = [a for a in sys.argv[1:] if a.startswith("-D")]
    if defs:
        print(build(force=True, verbose="--verbose" in sys.argv, defines=defs, out=HERE / "libb2m_alt.so"))
    else:
        print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
