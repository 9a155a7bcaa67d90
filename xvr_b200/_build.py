"""Build libxvr_b200.so (hand-written sm_100a kernels + the C-ABI of include/xvr_b200.h) in-tree with nvcc."""

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = Path(os.environ["XVR_B200_LIB"]).resolve() if os.environ.get("XVR_B200_LIB") else PKG / "libxvr_b200.so"  # tuning variants
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# siddon*.cu must reproduce the reference's un-fused fp32 arithmetic bit for bit -> no FMA contraction there
PER_FILE_FLAGS = {"siddon.cu": ["-fmad=false"], "siddon_volgrad.cu": ["-fmad=false"]}


def sources():
    return sorted(CSRC.glob("*.cu"))


def needs_build():
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ for sm_100a and link one shared library next to this file."""
    if os.environ.get("XVR_B200_LIB") or (not force and not needs_build()):
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build xvr_b200/libxvr_b200.so")
    objdir = PKG / "build"
    objdir.mkdir(exist_ok=True)
    procs = []
    for src in sources():
        obj = objdir / (src.stem + ".o")
        cmd = [nvcc, *ARCH, "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
               *PER_FILE_FLAGS.get(src.name, []), *os.environ.get("XVR_B200_NVCC_FLAGS", "").split(),
               "-c", str(src), "-o", str(obj)]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    log = []
    for src, obj, p in procs:
        out, _ = p.communicate()
        log.append(f"== {src.name}\n{out}")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src.name}:\n{out}")
        objs.append(str(obj))
    (objdir / "ptxas.log").write_text("\n".join(log))
    if verbose:
        print("\n".join(log))
    tmp = str(LIB) + ".tmp"
    cmd = [nvcc, *ARCH, "-shared", "-o", tmp, *objs]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose=True))
