"""Compare the SASS of every kernel in two builds of libs3g_b200.so (e.g. the last GPU-validated commit vs HEAD):

    git worktree add /tmp/val <validated-commit> && (cd /tmp/val && python -m s3gaussian_b200.build)
    python tools/compare_sass.py /tmp/val/s3gaussian_b200/lib/libs3g_b200.so s3gaussian_b200/lib/libs3g_b200.so

Used at the end of round 1, when code was added without GPU budget left, to prove that every kernel on a default
path was byte-identical to the binary the GPU tests had validated."""
import hashlib
import re
import subprocess
import sys


def kernels(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    d, cur, buf = {}, None, []
    for line in out.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if cur:
                d.setdefault(cur, []).append(hashlib.md5("\n".join(buf).encode()).hexdigest())
            cur, buf = m.group(1), []
        elif re.match(r"\s+/\*[0-9a-f]{4}\*/", line):
            buf.append(re.sub(r"/\*[0-9a-f]*\*/", "", line).strip())
    if cur:
        d.setdefault(cur, []).append(hashlib.md5("\n".join(buf).encode()).hexdigest())
    return d


def main():
    a, b = kernels(sys.argv[1]), kernels(sys.argv[2])
    same = [k for k in a if k in b and set(a[k]) <= set(b[k])]
    diff = [k for k in a if k in b and not set(a[k]) <= set(b[k])]
    print(f"{len(a)} kernels in {sys.argv[1]}: {len(same)} identical in {sys.argv[2]}, {len(diff)} differ, "
          f"{len([k for k in a if k not in b])} missing, {len([k for k in b if k not in a])} new")
    for k in diff:
        print("DIFF", k)
    for k in a:
        if k not in b:
            print("GONE", k)
    for k in b:
        if k not in a:
            print("NEW ", k)
    return 1 if diff else 0


if __name__ == "__main__":
    sys.exit(main())
