"""The C-ABI library loads in the GPU-less container and exports exactly the symbols declared in
include/polydis_b200.h, each bound in the ctypes table with the declared arity (no compute calls)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "polydis_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\bint\s+(pd_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        out[m.group(1)] = len([a for a in m.group(2).split(",") if a.strip()])
    return out


def test_header_declares_what_the_library_exports():
    from polydis_b200 import build
    lib = build.build()
    decl = _declared()
    assert len(decl) >= 25
    nm = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in nm.splitlines() if " T pd_" in l}
    assert exported == set(decl), (sorted(exported - set(decl)), sorted(set(decl) - exported))


def test_ctypes_table_matches_header():
    from polydis_b200 import _lib
    decl = _declared()
    assert set(_lib.SIGNATURES) == set(decl)
    for name, n_args in decl.items():
        assert len(_lib.SIGNATURES[name]) == n_args, name
        assert getattr(_lib.lib, name).restype is ctypes.c_int


def test_no_cpu_fallback():
    """The product refuses host tensors instead of silently computing on the CPU."""
    import torch
    from polydis_b200 import ops
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.linear(torch.zeros(2, 8), torch.zeros(4, 8), None)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.grid_prepare(torch.zeros(1, 32, 16, 6, dtype=torch.int64))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "polyphonic-chord-texture-disentanglement_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, fn)).read().replace("the CPU oracle", ""), fn
