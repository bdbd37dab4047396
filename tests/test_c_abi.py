"""CPU: the C-ABI libraries load and export every symbol include/*.h declares; with no GPU the
product fails loudly instead of falling back."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(agpt_[a-z0-9_]+)\s*\(", text)))


def test_core_exports_every_declared_symbol(agpt):
    lib = agpt.core()
    names = declared("agpt.h")
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_host_exports_every_declared_symbol(agpt):
    lib = agpt.host()
    names = [n for n in declared("agpt_host.h") if n.startswith("agpt_host_")]
    assert len(names) >= 15
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_struct_layouts(agpt):
    assert ctypes.sizeof(agpt.Material) == 64
    assert agpt.HIT_DTYPE.itemsize == 16
    assert ctypes.sizeof(agpt.Stats) == 7 * 8 + 6 * 16 + 5 * 8 + 6 * 4 + 4 + 4     # + reduce_path, trailing pad to 8


def test_no_cpu_fallback(agpt):
    """Without a CUDA device agpt_create must fail with a message; nothing renders on the CPU."""
    if agpt.device_count() > 0:
        pytest.skip("a GPU is visible here")
    with pytest.raises(agpt.AgptError) as e:
        agpt.Context(0)
    assert "no CPU path" in str(e.value) or "CUDA" in str(e.value)
    with pytest.raises(agpt.AgptError):
        agpt.HostTracer(5, 0)


def test_product_does_not_reference_the_oracle():
    """The shipped path must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "ag-pathtracer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".h", ".cpp", ".cu", ".cuh")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="replace").read()
                code = "\n".join(line for line in text.splitlines() if not line.strip().startswith(("//", "#", "*", "/*")))
                assert "libagpt_oracle" not in code and "libagpt_ref" not in code and "ref_binding" not in code, os.path.join(dirpath, f)
    import subprocess
    out = subprocess.run(["ldd", os.path.join(pkg, "libagpt_host.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out
