"""In-tree build of libsplat_b200.so (sm_100a only) with plain nvcc — no torch types in the library.

`python -m splatfields_b200.build` or `__graft_entry__.build()`.  The .so lands next to the sources
(git-ignored, but it travels to the GPU box with the gpurun snapshot).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(CSRC, "build")
LIB = os.path.join(HERE, "libsplat_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
# per-file extra flags: the preprocess kernel's float math feeds integer tile keys and must be
# bit-reproducible against the CPU oracle -> no implicit FMA contraction there.
EXTRA = {"preprocess.cu": ["-fmad=false"]}
SOURCES = ["api.cu", "preprocess.cu", "binning.cu", "render_fwd.cu", "render_bwd.cu", "geom_bwd.cu", "exchange.cu",
           "loss.cu", "densify.cu", "activate.cu", "knn.cu"]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(out: str, deps) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


# Build variants (A/B experiments; the default library is the product): name -> (extra nvcc flags, library path)
VARIANTS = {
    "": ([], LIB),
    "exactexp": (["-DSFB_EXACT_EXP"], os.path.join(HERE, "libsplat_b200_exactexp.so")),
}


def build(force: bool = False, verbose: bool = False, variant: str = "", force_sources=()) -> str:
    vflags, LIB = VARIANTS[variant]
    BUILD = os.path.join(CSRC, "build" + ("_" + variant if variant else ""))
    os.makedirs(BUILD, exist_ok=True)
    hdrs = [os.path.join(CSRC, "common.cuh"), os.path.join(HERE, "..", "include", "splat_b200.h"), __file__]
    objs = [os.path.join(BUILD, src.replace(".cu", ".o")) for src in SOURCES]

    def compile_one(src):
        sp = os.path.join(CSRC, src)
        op = os.path.join(BUILD, src.replace(".cu", ".o"))
        if not (force or src in force_sources or _stale(op, [sp] + hdrs)):
            return
        cmd = [nvcc()] + ARCH + COMMON + EXTRA.get(src, []) + vflags + ["-c", sp, "-o", op]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = os.path.join(BUILD, src + ".ptxas.log")
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")

    # translation units are independent: compile them side by side
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        list(ex.map(compile_one, SOURCES))
    force = force or bool(force_sources)
    if force or _stale(LIB, objs):
        cmd = [nvcc()] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart"]
        subprocess.check_call(cmd)
    return LIB


MANIFEST = os.path.join(HERE, "sass_manifest.json")


def sass_digest(lib: str = LIB) -> dict:
    """sha256 over the SASS of every kernel in the library (mangled name + instruction text and operands, addresses and
    encodings stripped): identifies the machine code independently of link order and file timestamps."""
    import hashlib
    import re
    cuobjdump = os.path.join(os.path.dirname(nvcc()), "cuobjdump")
    txt = subprocess.run([cuobjdump, "-sass", lib], capture_output=True, text=True, check=True).stdout
    kernels, cur = {}, None
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            kernels[cur] = hashlib.sha256()
        elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
            ins = re.sub(r"\s+", " ", re.sub(r"/\*[0-9a-f]+\*/", "", ln)).strip()
            kernels[cur].update(ins.encode() + b"\n")
    h = hashlib.sha256()
    for name in sorted(kernels):
        h.update(name.encode() + b"=" + kernels[name].hexdigest().encode() + b"\n")
    ver = subprocess.run([nvcc(), "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-1]
    return {"digest": h.hexdigest(), "kernels": len(kernels), "nvcc": ver}


def write_manifest() -> dict:
    import json
    d = sass_digest(build())
    with open(MANIFEST, "w") as f:
        json.dump(d, f, indent=1)
        f.write("\n")
    return d


def check_manifest() -> dict:
    """The committed manifest must describe the library that was just built from the committed sources."""
    import json
    got = sass_digest(LIB)
    want = json.load(open(MANIFEST))
    if got["digest"] != want["digest"] or got["kernels"] != want["kernels"]:
        raise RuntimeError(f"libsplat_b200.so does not match splatfields_b200/sass_manifest.json: built {got}, "
                           f"manifest {want} (after changing a kernel: python -m splatfields_b200.build --write-manifest)")
    return got


if __name__ == "__main__":
    if "--write-manifest" in sys.argv:
        print(write_manifest())
        sys.exit(0)
    var = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--variant=")), "")
    print(build(force="--force" in sys.argv, verbose="--quiet" not in sys.argv, variant=var))
