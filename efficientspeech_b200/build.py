"""Build libes_b200.so (sm_100a only) in-tree with nvcc.  No torch headers are involved: the
library is a plain C-ABI shared object (include/es_b200.h)."""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libes_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
    "-Xptxas", "-v", "--expt-relaxed-constexpr",
]


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.isfile(cand):
            return cand
    raise RuntimeError("nvcc not found (set $NVCC)")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build() -> bool:
    if not os.path.isfile(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        [os.path.join(os.path.dirname(PKG_DIR), "include", "es_b200.h"), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), out: str = LIB_PATH) -> str:
    """One object per .cu (compiled in parallel, rebuilt only when the source or a header changed), then one link.
    `defines` / `out`: experiment builds (e.g. -DES_MBAR_SUSPEND_NS=0 into a second .so selected with $ES_B200_LIB);
    their objects live in their own directory.  The default build takes neither."""
    if not force and not defines and out == LIB_PATH and not needs_build():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor
    nvcc = find_nvcc()
    tag = "default" if not defines else "exp_" + "_".join(d.replace("=", "-") for d in defines)
    objdir = os.path.join(PKG_DIR, "build", tag)
    os.makedirs(objdir, exist_ok=True)
    headers = glob.glob(os.path.join(CSRC, "*.cuh")) + \
        [os.path.join(os.path.dirname(PKG_DIR), "include", "es_b200.h"), os.path.abspath(__file__)]
    hdr_t = max(os.path.getmtime(h) for h in headers)
    cflags = [f for f in NVCC_FLAGS if f != "-shared"] + [f"-D{d}" for d in defines]

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.isfile(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t):
            return obj, 0, ""
        cmd = [nvcc] + cflags + ["-c", src, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return obj, res.returncode, " ".join(cmd) + "\n" + res.stdout + res.stderr

    with ThreadPoolExecutor(max(1, min(8, os.cpu_count() or 1))) as ex:
        results = list(ex.map(compile_one, sources()))
    log = "".join(r[2] for r in results)
    ok = all(r[1] == 0 for r in results)
    if ok:
        cmd = [nvcc, "-shared", "-Xcompiler", "-fPIC"] + [r[0] for r in results] + ["-o", out, "-lcudart"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        log += " ".join(cmd) + "\n" + res.stdout + res.stderr
        ok = res.returncode == 0
    if not defines:
        with open(os.path.join(PKG_DIR, "build.log"), "a" if not force else "w") as f:
            f.write(log)
    if not ok:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libes_b200.so")
    if verbose:
        print(log)
    return out


if __name__ == "__main__":
    defs = [a[2:] for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[6:] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose=not defs, defines=defs, out=outs[0] if outs else LIB_PATH))
