"""CPU: the C-ABI shared library loads and exports every symbol include/auvpf.h
declares; no compute is called (no GPU here)."""
import os
import re

import pytest

from smarc_navigation_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "auvpf.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(auvpf_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    names = _declared()
    assert len(names) >= 30
    lib = _lib.load()
    bound = {n for n, _, _ in _lib.SYMBOLS}
    for n in names:
        assert hasattr(lib, n), "libauvpf.so does not export %s" % n
        assert n in bound, "%s declared in auvpf.h but not bound in _lib.SYMBOLS" % n
    assert bound <= set(names)
    assert lib.auvpf_version() == 202


def test_config_struct_matches_header():
    import ctypes as C
    # int32 x2, int64, int32 x2, int64 x2, double[6], double, double[16], uint64, double x2
    assert C.sizeof(_lib.Config) == 4 + 4 + 8 + 4 + 4 + 8 + 8 + 48 + 8 + 128 + 8 + 8 + 8


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from smarc_navigation_b200 import ParticleFilter, AuvpfError
    with pytest.raises(AuvpfError):
        ParticleFilter(particle_count=16)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "smarc_navigation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "oracle/" not in txt.replace(
                    "oracle/c/auvpf_oracle.c RAYCAST SPEC", "").replace("oracle/pf_oracle.py quantise_logw", ""), f


def test_parse_covariance_strings():
    from smarc_navigation_b200 import parse_covariance
    assert parse_covariance("[0.1, 0.1, 0.0, 0.0, 0.0, 0.0]") == [0.1, 0.1, 0, 0, 0, 0]        # auv_pf.launch:17
    assert parse_covariance("[1., 1., 0.0, 0.0, 0.0, 0.0001]")[5] == 1e-4                        # auv_pf.launch:19
    with pytest.raises(ValueError):
        parse_covariance([1, 2, 3])
