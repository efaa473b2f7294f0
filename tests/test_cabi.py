"""The C-ABI library loads and exports every symbol include/td_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    src = open(os.path.join(ROOT, "include", "td_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(td_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from multidiffusion_upscaler_for_automatic1111_b200 import _cabi
    declared = _declared()
    assert len(declared) >= 10
    raw = ctypes.CDLL(_cabi.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in td_b200.h but not exported by libtd_b200.so"
    assert sorted(_cabi.exported_symbols()) == declared, "ctypes binding and header disagree"


def test_abi_version_and_error_string():
    from multidiffusion_upscaler_for_automatic1111_b200 import _cabi
    assert _cabi.lib.td_abi_version() == _cabi.ABI_VERSION
    st = _cabi.lib.td_split_bboxes(64, 64, 16, 16, 16, None, 0, None, None)  # overlap == tile: invalid
    assert st == _cabi.TD_ERR_INVALID_ARG
    assert b"overlap" in _cabi.lib.td_last_error()


def test_struct_layout_matches_header():
    from multidiffusion_upscaler_for_automatic1111_b200 import _cabi
    assert ctypes.sizeof(_cabi.TdGrid) == 4 * (10 + 2 * _cabi.TD_MAX_GRID_DIM)


def test_missing_library_fails_loudly(tmp_path, monkeypatch):
    """No silent fallback: a missing .so is an ImportError naming the build command."""
    import importlib

    from multidiffusion_upscaler_for_automatic1111_b200 import _cabi
    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    try:
        _cabi._load()
    except ImportError as e:
        assert "no CPU" in str(e) or "fallback" in str(e)
    else:
        raise AssertionError("loading a missing library must raise")
    importlib.reload  # keep linters quiet
