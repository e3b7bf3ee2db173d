"""CPU test: the C-ABI library loads without a GPU and exports every symbol
that include/hydrob200.h declares (no compute calls here)."""
import ctypes as ct
import os
import re

from conftest import ROOT
from hydromodel_b200 import capi


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "hydrob200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hb_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree():
    names = declared_symbols()
    assert len(names) >= 25
    assert set(names) == set(capi.EXPORTS), set(names) ^ set(capi.EXPORTS)


def test_library_loads_and_exports_all_symbols():
    lib = capi.load()
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.hb_version() >= 100
    assert lib.hb_max_nx() >= 1001


def test_no_cpu_fallback_without_device():
    """Without a CUDA device hb_create must fail with a message, not fall back."""
    try:
        import torch
        if torch.cuda.is_available():
            return
    except Exception:      # noqa: BLE001
        pass
    lib = capi.load()
    h = lib.hb_create(0, 401, 1)
    assert not h
    assert b"no CUDA device" in lib.hb_last_error() or b"cuda" in lib.hb_last_error().lower()


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under hydromodel_b200/ may
    import or load it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "hydromodel_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(import|from)\s+oracle\b", text, flags=re.M) or "liboracle" in text:
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
