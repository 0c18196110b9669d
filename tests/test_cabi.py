"""The C-ABI library: loads without a GPU, exports every symbol include/poi_engine.h declares, has
no torch symbols in its interface, and fails loudly (no CPU fallback) when there is no device."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "poi_engine.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(poi_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    import poi_b200  # noqa: F401
    from poi_b200 import _lib
    assert _declared() == _lib.EXPORTED


def test_library_exports_every_declared_symbol():
    import poi_b200  # noqa: F401
    from poi_b200 import _lib
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = set(l.split()[-1] for l in out.splitlines() if " T " in l)
    missing = [s for s in _declared() if s not in exported]
    assert not missing, missing
    # C linkage only: no mangled torch / at:: / c10:: symbols leak into the ABI
    assert not [s for s in exported if "torch" in s or "c10" in s or "at::" in s]


def test_no_cpu_fallback():
    import torch
    import poi_b200  # noqa: F401
    from poi_b200.engine import Engine, EngineError
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(EngineError):
        Engine(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "point-of-interest-recommendation_b200")
    bad = []
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M):
                    bad.append(f)
    assert not bad, bad
