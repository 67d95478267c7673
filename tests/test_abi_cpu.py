"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/colord_b200.h declares,
refuses to compute without a GPU (no CPU fallback), and its host-side sampler matches the oracle + goldens."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle_lib
from colord_b200 import lib
from conftest import GOLDEN_CASES, ROOT


def _declared():
    with open(os.path.join(ROOT, "include", "colord_b200.h")) as f:
        src = f.read()
    return sorted(set(re.findall(r"\b(clb_[a-z0-9_]+)\s*\(", src)))


def test_exports_every_declared_symbol():
    L = ctypes.CDLL(lib.SO_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), n
    assert set(names) == set(lib.EXPORTS)


def test_mgpu_library_exports_every_declared_symbol():
    """include/colord_b200_mgpu.h <-> libcolord_b200_mgpu.so (the NCCL exchanges of the multi-GPU path)"""
    with open(os.path.join(ROOT, "include", "colord_b200_mgpu.h")) as f:
        names = sorted(set(re.findall(r"\b(clb_group_[a-z0-9_]+)\s*\(", f.read())))
    assert len(names) == 5
    # read with nm, not loaded: the library brings the system's NCCL with it, and a process that imports torch afterwards would
    # bind torch to that copy instead of its own
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(ROOT, "colord_b200", "libcolord_b200_mgpu.so")], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (clb_group_[a-z0-9_]+)", out))
    assert exported == set(names)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lib.ClbError) as e:
        lib.Context(20, 12, 4, 80, 5)
    assert e.value.status == 1


def test_bad_params_rejected():
    for args in [(40, 12, 4, 80, 5), (20, 0, 4, 80, 5), (20, 12, 4, 2, 5), (20, 12, 4, 80, 0)]:
        with pytest.raises(lib.ClbError) as e:
            lib.Context(*args)
        assert e.value.status in (1, 3)


@pytest.mark.parametrize("rng,exp,n_pseudo,n", [(1, 1.0, 0, 1000), (7, 1.0, 0, 5000), (46, 1.0, 3, 3000), (5, 0.5, 0, 2000), (3, 2.0, 10, 500)])
def test_sampler_matches_oracle(rng, exp, n_pseudo, n):
    assert np.array_equal(lib.sampler(rng, exp, n_pseudo, n), oracle_lib.sampler(rng, exp, n_pseudo, n))


@pytest.mark.parametrize("case", GOLDEN_CASES)
def test_sampler_matches_reference(golden, case):
    g = golden(case)
    if not g.params["sparse"]:
        pytest.skip("-R all")
    dec = lib.sampler(g.params["sparse_range"], float(g.params["sparse_exponent"]), 0, len(g.reads))
    assert np.array_equal(dec & (1 - g.has_n), g.is_ref)


def test_cpp_host_mirror_compiles():
    """The C++ host-side mirror of the reference classes (colord_b200/host) builds against the C-ABI."""
    import subprocess
    src = '#include "colord_b200/host/stage23_host.h"\nint main() { clbhost::CRefReadsAccepter a(7, 1.0, 0); return a.GetNAccepted(100) > 100; }\n'
    subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I", ROOT, "-x", "c++", "-"], input=src.encode(), check=True, cwd=ROOT)


def test_slab_bookkeeping(tmp_path):
    """colord_b200/csrc/slab.h (the optional per-process device slab, CLB_SLAB_GB) is pure host bookkeeping: random
    allocate / free sequences keep its invariants (alignment, no overlap, merging of freed ranges, refusal only when nothing fits)."""
    import subprocess
    exe = str(tmp_path / "slabtest")
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-o", exe, os.path.join(ROOT, "tests", "host_slab_test.cpp")], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "slab ok" in r.stdout, r.stderr
