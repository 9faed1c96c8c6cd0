"""The C-ABI library loads without a GPU and exports every symbol include/mmesh_b200.h declares (no compute calls)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exports_match_header():
    from mmesh_b200 import capi
    so = capi.build()
    L = ctypes.CDLL(so)
    hdr = open(os.path.join(ROOT, "include", "mmesh_b200.h")).read()
    declared = set(re.findall(r"\b(mmb_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"mmb_context", "mmb_status", "mmb_params", "mmb_step_result"}
    assert len(declared) > 30
    for name in sorted(declared):
        assert hasattr(L, name), name
    assert declared == set(capi.EXPORTS), declared ^ set(capi.EXPORTS)


def test_no_gpu_is_an_error_not_a_fallback():
    """Without a usable device mmb_create must fail (status != 0): the product path never computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        return
    from mmesh_b200 import capi
    import pytest
    with pytest.raises(capi.MMeshError):
        capi.Context(dim=2, device=0)


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "dune-mmesh_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hh", ".cc")):
                txt = open(os.path.join(base, f), errors="ignore").read()
                assert "pyoracle" not in txt and "mmesh_oracle" not in txt and "orc_" not in txt, os.path.join(base, f)
