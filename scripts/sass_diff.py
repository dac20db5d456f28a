"""Which kernels changed since <git-rev>?  Compiles that revision's csrc/*.cu next to the current build and compares the
SASS of every kernel present in both, instruction text and operands included.  Used to show that opt-in variants (added
as separate instantiations / wrappers) leave the default kernels bit-identical to a build that ran on the GPU.
    python scripts/sass_diff.py 34bb199"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from splatfields_b200 import build as B   # noqa: E402


def kernels(obj):
    out, cur = {}, None
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            out[cur] = []
        elif cur and re.match(r"\s+/\*[0-9a-f]{4}\*/", ln):
            out[cur].append(re.sub(r"\s+", " ", re.sub(r"/\*[0-9a-f]+\*/", "", ln)).strip())
    return out


def main(rev):
    B.build()
    same = changed = 0
    only_new = []
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(os.path.join(td, "splatfields_b200", "csrc"))
        os.makedirs(os.path.join(td, "include"))
        files = subprocess.run(["git", "ls-tree", "--name-only", rev, "splatfields_b200/csrc/"], cwd=ROOT,
                               capture_output=True, text=True).stdout.split()
        for f in files + ["include/splat_b200.h"]:
            if f.endswith((".cu", ".cuh", ".h")):
                open(os.path.join(td, f), "w").write(
                    subprocess.run(["git", "show", f"{rev}:{f}"], cwd=ROOT, capture_output=True, text=True).stdout)
        for f in files:
            if not f.endswith(".cu"):
                continue
            src = os.path.basename(f)
            cur_obj = os.path.join(B.CSRC, "build", src.replace(".cu", ".o"))
            if not os.path.exists(cur_obj):
                continue
            old_obj = os.path.join(td, src + ".o")
            subprocess.check_call([B.nvcc()] + B.ARCH + ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC"] +
                                  B.EXTRA.get(src, []) + ["-c", src, "-o", old_obj],
                                  cwd=os.path.join(td, "splatfields_b200", "csrc"))
            old, new = kernels(old_obj), kernels(cur_obj)
            # exact mangled name first; a kernel that gained defaulted template parameters has a longer argument list,
            # so fall back to "same function stem and identical body"
            for name, body in new.items():
                match = name if name in old else None
                if match is None:
                    stem = re.match(r"(_ZN3sfb\d+\w+?I)", name)
                    match = next((o for o in old if stem and o.startswith(stem.group(1)) and old[o] == body), None)
                if match is None:
                    only_new.append(name)
                elif old[match] == body:
                    same += 1
                else:
                    changed += 1
                    print("CHANGED", name)
    print(f"{same} kernels identical to {rev}, {changed} changed, {len(only_new)} without a counterpart (new or altered):")
    for n in only_new:
        print("   ", subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()[:140])


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "HEAD")
