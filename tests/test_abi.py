"""The C-ABI library loads and exports every symbol include/defuse_b200.h declares.  No compute
calls: this runs on the CPU-only box.  Also checks that the product fails loudly without a GPU."""
import ctypes
import os
import re

import pytest

import defuse_b200 as d

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "defuse_b200.h")).read()
    return sorted(set(re.findall(r"DFB_API\s+[\w\s\*]+?\b(dfb_\w+)\s*\(", text)))


def test_library_is_built():
    assert os.path.exists(d.LIB_PATH), "run __graft_entry__.build() first"


def test_every_declared_symbol_is_exported():
    names = _declared()
    assert len(names) >= 24
    lib = ctypes.CDLL(d.LIB_PATH)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    # and the Python binding knows each of them
    assert sorted(d.ABI_SYMBOLS) == names


def test_abi_version_and_struct_sizes():
    lib = d.load_library()
    assert lib.dfb_abi_version() == 1
    assert d.SPLIT_ROW_DTYPE.itemsize == 32
    assert ctypes.sizeof(d._SeqTable) == 24
    assert ctypes.sizeof(d._PlanStats) == 11 * 8 + 3 * 8


def test_library_has_sm100a_code_only():
    out = os.popen("cuobjdump -lelf %s 2>/dev/null" % d.LIB_PATH).read()
    if not out:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(d.DefuseB200Error) as e:
        d.Context(0)
    assert e.value.status == 4  # DFB_ERR_NODEVICE
    assert "no CPU fallback" in str(e.value) or "device" in str(e.value)


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "defuse_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "import oracle" not in text and "dp_oracle" not in text and "oracle/" not in text, os.path.join(dirpath, f)
