"""CPU-side checks of the boundary: the library builds/loads here (nvcc cross-compiles without a GPU), exports every
symbol declared in include/softmac_b200.h, and refuses to run without a device instead of falling back."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "softmac_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(smx_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    from softmac_b200 import build as b
    from softmac_b200 import _capi
    b.build()
    lib = ctypes.CDLL(_capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/softmac_b200.h but not exported"
    assert sorted(_capi.EXPORTED_SYMBOLS) == names, "ctypes binding table and header disagree"


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from softmac_b200._capi import SmxError
    from softmac_b200.engine import MPMSimulator
    from softmac_b200.config import simulator_defaults
    cfg = simulator_defaults()
    cfg.n_particles = 16
    with pytest.raises(SmxError) as e:
        MPMSimulator(cfg)
    assert "no CPU path" in str(e.value)


def test_product_never_imports_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle[/.]mpm_oracle|liboracle", re.M)
    for dp, _, fs in os.walk(os.path.join(ROOT, "softmac_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not pat.search(txt), f"{os.path.join(dp, f)} references the oracle"
