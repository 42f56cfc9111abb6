"""gcc build of the oracle's multi-threaded C restatement (oracle/simples_oracle_c.c) -> oracle/libsimples_oracle_c.so.
TEST / BASELINE INFRASTRUCTURE: built by __graft_entry__.build() next to the CUDA library; the .so is git-ignored and
travels to the GPU box with the snapshot."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "simples_oracle_c.c")
LIB = os.path.join(HERE, "libsimples_oracle_c.so")


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        # -O2 without -march / -ffast-math: the box's host CPU may differ from the build container's, and the arithmetic
        # must be the plain IEEE evaluation of the formulas
        # some images wrap gcc so that libgomp.spec is not on the spec search path: retry with the compiler's install
        # directory as -B, then with the system gcc; without OpenMP at all the library still builds (one thread) and
        # so_max_threads() says so
        tries = []
        for cc in dict.fromkeys([os.environ.get("CC", "gcc"), "gcc", "/usr/bin/gcc"]):
            try:
                inst = subprocess.run([cc, "-print-search-dirs"], capture_output=True, text=True).stdout.split("\n")[0]
            except OSError:
                continue
            inst = inst.split("install:")[-1].strip()
            tries += [(cc, ["-fopenmp"]), (cc, ["-fopenmp", "-B" + inst])]
        tries.append(("gcc", []))
        for cc, extra in tries:
            r = subprocess.run([cc, "-O2", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"] + extra, capture_output=True, text=True)
            if r.returncode == 0:
                break
        else:
            raise RuntimeError(f"gcc failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force=True))
