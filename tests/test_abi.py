"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/*.h
declares, validates arguments, and FAILS LOUDLY (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "simples_b200.h")) as f:
        src = f.read()
    return sorted(set(re.findall(r"SIMPLES_API[^;(]*?\b(simples_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built_lib):
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(built_lib, s), f"{s} declared in include/simples_b200.h but not exported"
    from simples_b200 import _lib
    assert sorted(_lib.EXPORTS) == syms
    assert built_lib.simples_abi_version() == 1


def test_kernel_names(built_lib):
    n = built_lib.simples_kernel_count()
    names = [built_lib.simples_kernel_name(i).decode() for i in range(n)]
    assert "estep_project" in names and "impute_sweep" in names and "mstep_stats" in names and len(set(names)) == n


def test_struct_layout_matches_header():
    """ctypes mirrors of the POD structs: field order and sizes follow the header."""
    from simples_b200 import _lib
    assert C.sizeof(_lib.EmOut) == 13 * 8 and C.sizeof(_lib.MiOut) == 6 * 8
    assert _lib.EmArgs.G.offset == 0 and _lib.EmArgs.Y.offset == 24 and _lib.EmArgs.cutoff.offset == 48
    assert _lib.EmArgs.seed.offset % 8 == 0 and C.sizeof(_lib.EmArgs) % 8 == 0
    assert _lib.MiArgs.dat.offset == 24


def test_struct_layout_against_gcc(tmp_path):
    """sizeof / offsetof of every struct as gcc lays out include/simples_b200.h == the ctypes mirrors."""
    import os, subprocess
    from simples_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pairs = [("simples_em_args", _lib.EmArgs), ("simples_em_out", _lib.EmOut), ("simples_mi_args", _lib.MiArgs),
             ("simples_mi_out", _lib.MiOut), ("simples_lq_args", _lib.LqArgs), ("simples_lq_out", _lib.LqOut),
             ("simples_init_args", _lib.InitArgs), ("simples_init_out", _lib.InitOut),
             ("simples_clus_args", _lib.ClusArgs), ("simples_clus_out", _lib.ClusOut)]
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "simples_b200.h"', 'int main(void) {']
    for cname, cls in pairs:
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname.rstrip("_")}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in pairs:
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, (cname, fname)


def test_argument_validation_without_gpu(built_lib):
    import simples_b200 as sb
    G, n, K0, M0 = 4, 3, 2, 1
    ok = dict(Y=np.ones((G, n)), Y0=np.ones((G, n)), pg=np.ones((G, 1)), beta=np.ones((G, K0)), sigma=np.ones((G, M0)),
              lam=np.ones((M0, K0)), pi=[1.0])
    with pytest.raises(sb.SimplesError) as e:       # K0 out of range -> EINVAL before any device work
        sb.EM_impute(ok["Y"], ok["Y0"], ok["pg"], M0, 40, 0.1, 1, np.ones((G, 40)), ok["sigma"], np.ones((M0, 40)), ok["pi"], None)
    assert e.value.code == -1
    with pytest.raises(ValueError):                  # host-side shape check
        sb.EM_impute(ok["Y"], ok["Y0"][:, :2], ok["pg"], M0, K0, 0.1, 1, ok["beta"], ok["sigma"], ok["lam"], ok["pi"], None)


def test_no_cpu_fallback(built_lib):
    """Without a CUDA device the product path must raise, not compute."""
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    import simples_b200 as sb
    G, n, K0, M0 = 4, 3, 2, 1
    with pytest.raises(sb.SimplesError) as e:
        sb.EM_impute(np.ones((G, n)), np.ones((G, n)), np.ones((G, 1)), M0, K0, 0.1, 1, np.ones((G, K0)), np.ones((G, M0)),
                     np.ones((M0, K0)), [1.0], None)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    with pytest.raises(sb.SimplesError):
        sb.do_impute(np.ones((G, n)), np.ones((G, n)), np.ones((G, K0)), np.ones((M0, K0)), np.ones((G, M0)),
                     np.zeros((G, M0)), [1.0])


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "simples_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                with open(os.path.join(dp, f)) as fh:
                    src = fh.read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), f


def test_host_mask_packing_matches_numpy(built_lib):
    """simples_pack_mask (host threads + SSE2/AVX2 compares, no GPU): bit j of word [w][i] <=> Y0[32 w + j, i] <= cutoff,
    including NaN (never masked), +-0, values exactly at the cutoff, and a ragged last word."""
    import ctypes as C
    import time
    import numpy as np
    rng = np.random.default_rng(0)
    for G, n in ((2000, 3000), (77, 513), (32, 1), (1, 300)):
        Y0 = np.asfortranarray(np.where(rng.random((G, n)) < 0.45, 0.0, rng.gamma(2.0, 1.0, (G, n))))
        Y0[rng.integers(0, G, 20), rng.integers(0, n, 20)] = 0.1            # exactly at the cutoff: masked (<=)
        Y0[rng.integers(0, G, 5), rng.integers(0, n, 5)] = np.nan           # NaN: not masked
        Y0[rng.integers(0, G, 5), rng.integers(0, n, 5)] = -0.0
        nw = (G + 31) // 32
        out = np.zeros((nw, n), dtype=np.uint32)
        t0 = time.perf_counter()
        cnt = built_lib.simples_pack_mask(Y0.ctypes.data_as(C.c_void_p), G, n, 0.1, out.ctypes.data_as(C.c_void_p))
        dt = time.perf_counter() - t0
        with np.errstate(invalid="ignore"):
            bits = (Y0 <= 0.1)
        want = np.zeros((nw * 32, n), dtype=bool)
        want[:G] = bits
        w = (want.reshape(nw, 32, n).astype(np.uint64) << np.arange(32, dtype=np.uint64)[None, :, None]).sum(axis=1).astype(np.uint32)
        assert cnt == int(bits.sum()) and np.array_equal(out, w), (G, n)
    assert dt < 5.0
