"""ORACLE build recipe (test infrastructure): compiles oracle/csrc/segtree.c into
oracle/_build/liboracle_segtree.so with gcc.  No reference sources are compiled — the reference
is pure Python, so there is no ``oracle/_ref`` binary for this path."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "csrc", "segtree.c")
    out_dir = os.path.join(_HERE, "_build")
    out = os.path.join(out_dir, "liboracle_segtree.so")
    os.makedirs(out_dir, exist_ok=True)
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC",
                               "-o", out, src, "-lm"])
    return out


if __name__ == "__main__":
    print(build(force=True))
