"""Per-kernel comparison of the device code of two builds of libperc_b200.so (cuobjdump -sass, encodings and
column padding dropped).  Used to check that adding an opt-in kernel variant leaves the code of the kernels that
have been validated on the GPU untouched:  python tools/sass_diff.py old.so new.so [old_name=new_name ...]
(a rename maps a kernel whose mangled name changed, e.g. when a template parameter was added)."""
import collections, subprocess, sys


def load(path):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    d, name = collections.defaultdict(list), None
    for line in out.splitlines():
        if "Function :" in line:
            name = line.split("Function :")[1].strip()
            continue
        line = line.split("/* 0x")[0]
        if name and line.strip():
            d[name].append(" ".join(line.split()))
    return d


if __name__ == "__main__":
    a, b = load(sys.argv[1]), load(sys.argv[2])
    ren = dict(x.split("=") for x in sys.argv[3:])
    changed = missing = 0
    for k, v in a.items():
        k2 = ren.get(k, k)
        if k2 not in b:
            missing += 1
            print("MISSING", k)
        elif v != b[k2]:
            changed += 1
            print("CHANGED", k)
    print("kernels in %s: %d, changed: %d, missing: %d; kernels in %s: %d" % (sys.argv[1], len(a), changed, missing, sys.argv[2], len(b)))
