"""C-ABI checks that need no GPU: the library builds/loads and exports every symbol the header declares."""
import os
import re

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "sdrgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sdr_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(capi):
    L = capi.lib()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/sdrgpu.h but not exported by libsdrgpu.so"
    assert sorted(capi.SYMBOLS) == declared


def test_version_and_no_torch_types(capi):
    assert b"sm_100a" in capi.lib().sdr_version()
    # the boundary is plain C: no torch / C++ types in the header
    text = open(os.path.join(ROOT, "include", "sdrgpu.h")).read()
    assert "torch" not in text and "std::" not in text and "at::" not in text


def test_product_never_touches_the_oracle():
    """the product path must not import, link or call anything under oracle/"""
    pkg = os.path.join(ROOT, "sdrainer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                src = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert "sdr_oracle" not in src and "libsdroracle" not in src and "from oracle" not in src \
                    and "import oracle" not in src, f


def test_sass_has_tma_bulk_copy(capi):
    """K1/K3 stage IQ blocks with cp.async.bulk: the SASS must show UBLKCP (B200_PROFILING.md)"""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        import pytest
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", capi.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UBLKCP" in sass
    assert "SYNCS" in sass  # mbarrier


def test_loaded_library_is_never_rebuilt_in_place(monkeypatch):
    """A stale libsdrgpu.so that this process has already loaded must not be rebuilt (a rebuilt file is a second copy
    for every later dlopen: the host mirror then launches kernels whose attributes no engine set -> invalid argument)"""
    from sdrainer_b200 import _build, capi
    capi.lib()
    assert _build.LOADED
    monkeypatch.setattr(_build, "is_stale", lambda: True)
    calls = []
    monkeypatch.setattr(_build.subprocess, "check_call", lambda *a, **k: calls.append(a))
    assert _build.build() == _build.LIB and not calls
