"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/wafer_b200.h declares, and refuses to run (loudly) without a B200.  No compute calls."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "wafer_b200.h")


@pytest.fixture(scope="module")
def capi():
    import __graft_entry__
    if not os.path.exists(os.path.join(ROOT, "wafer_b200", "libwafer_b200.so")):
        __graft_entry__.build()
    from wafer_b200 import _capi
    _capi.load()
    return _capi


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wafer_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(capi):
    declared = _declared_symbols()
    assert len(declared) >= 30
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (wafer_[a-z0-9_]+)", out))
    assert set(declared) <= exported, sorted(set(declared) - exported)
    assert set(declared) == set(capi.SYMBOLS), sorted(set(declared) ^ set(capi.SYMBOLS))


def test_library_is_sm100a_only_and_has_no_oracle_dependency(capi):
    out = subprocess.run(["cuobjdump", "--list-elf", capi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs
    needed = subprocess.run(["readelf", "-d", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in needed
    strings = subprocess.run(["strings", capi.LIB_PATH], capture_output=True, text=True).stdout
    assert "libwafer_oracle" not in strings


def test_struct_layouts_match_header(capi):
    assert C.sizeof(capi.Observables) == 32
    assert C.sizeof(capi.Record) == 8 + 16 + 32
    assert capi.Params.nx.offset == 0 and capi.Params.ext.offset == 24 and capi.Params.dn.offset == 32
    assert capi.Params.device.offset == 56 and capi.Params.nccl_id.offset == 72 and C.sizeof(capi.Params) == 88


@pytest.mark.skipif(os.path.exists("/dev/nvidiactl"), reason="a GPU is present")
def test_no_cpu_fallback(capi):
    import wafer_b200
    with pytest.raises(wafer_b200.WaferError) as ei:
        wafer_b200.Lattice((8, 8, 8))
    assert ei.value.status == 2  # WAFER_ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "wafer_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), os.path.join(dirpath, f)
