"""Prove that a change to csrc/ left the already-verified kernels untouched: compile every .cu of a base git revision and of
the working tree for sm_100a (no GPU needed), and compare the SASS of every kernel that exists in both, instruction for
instruction (addresses and encodings stripped, anonymous-namespace hashes normalised).  New template instantiations are
listed, not compared.  Used when opt-in kernel variants are added next to GPU-verified defaults.
  python scripts/sass_diff.py <base-rev>          e.g. python scripts/sass_diff.py 4cc9997"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join("isca-2025-lia_b200", "csrc")
NVCC = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC"]


def sass(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = re.sub(r"_ZN\d+_GLOBAL__N__[0-9a-f]+_\d+_\w+?_cu_[0-9a-f]{8}", "_ZN_ANON_", m.group(1))
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?)\s*/\* 0x[0-9a-f]+ \*/", line)
        if m and cur is not None:
            funcs[cur].append(m.group(1))
    return funcs


def build(srcdir, incdir, name, tmp, tag):
    obj = os.path.join(tmp, f"{tag}_{name}.o")
    subprocess.run(NVCC + ["-I", incdir, "-I", srcdir, "-c", os.path.join(srcdir, name), "-o", obj], check=True, capture_output=True)
    return obj


def main(rev):
    with tempfile.TemporaryDirectory() as tmp:
        base = os.path.join(tmp, "base")
        os.makedirs(os.path.join(base, CSRC))
        os.makedirs(os.path.join(base, "include"))
        files = subprocess.run(["git", "ls-tree", "-r", "--name-only", rev, CSRC, "include"], cwd=ROOT, capture_output=True, text=True,
                               check=True).stdout.split()
        for f in files:
            data = subprocess.run(["git", "show", f"{rev}:{f}"], cwd=ROOT, capture_output=True, check=True).stdout
            os.makedirs(os.path.dirname(os.path.join(base, f)), exist_ok=True)
            open(os.path.join(base, f), "wb").write(data)
        bad = 0
        for name in sorted(n for n in os.listdir(os.path.join(ROOT, CSRC)) if n.endswith(".cu")):
            if not os.path.exists(os.path.join(base, CSRC, name)):
                print(f"{name}: new file")
                continue
            old = sass(build(os.path.join(base, CSRC), os.path.join(base, "include"), name, tmp, "old"))
            new = sass(build(os.path.join(ROOT, CSRC), os.path.join(ROOT, "include"), name, tmp, "new"))
            same = [k for k in old if k in new and old[k] == new[k]]
            diff = [k for k in old if k in new and old[k] != new[k]]
            gone = [k for k in old if k not in new]
            added = [k for k in new if k not in old]
            renamed = {}                      # a kernel that only gained a template parameter: same instructions, new name
            for k in list(gone):
                twin = next((a for a in added if new[a] == old[k]), None)
                if twin is not None:
                    renamed[k] = twin
                    gone.remove(k)
                    added.remove(twin)
            bad += len(diff) + len(gone)
            print(f"{name}: {len(same)} kernels identical, {len(renamed)} renamed but identical, {len(diff)} changed, "
                  f"{len(gone)} removed, {len(added)} new")
            for k, t in renamed.items():
                print("   renamed ", k[:110], "->", t[:110])
            for k in diff + gone:
                print("   CHANGED " if k in diff else "   GONE    ", k[:150])
            for k in added:
                print("   new     ", k[:150])
        return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1] if len(sys.argv) > 1 else "HEAD"))
