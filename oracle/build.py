"""Build recipe for the CPU oracle shared library (test infrastructure).

Same flags as oracle/Makefile; falls back to a build without OpenMP when the
toolchain has no libgomp.  Output: oracle/libsurfel_oracle.so (git-ignored).
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "surfel_oracle.c")
OUT = os.path.join(HERE, "libsurfel_oracle.so")


def build(force: bool = False) -> str:
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= os.path.getmtime(SRC)):
        return OUT
    base = ["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
            "-Wall", "-Wno-unused-variable", "-Wno-unknown-pragmas"]
    last = None
    for extra in (["-fopenmp"], []):
        cmd = base + extra + ["-o", OUT, SRC, "-lm"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode == 0:
            return OUT
        last = r.stderr
    raise RuntimeError("oracle build failed:\n" + str(last))


if __name__ == "__main__":
    print(build(force=True))
