"""The C-ABI library loads without a GPU and exports every symbol include/jagua_cuda.h declares; the
product path fails loudly (no CPU fallback) when no CUDA device is usable."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "jagua_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(jg_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for must in ("jg_engine_create", "jg_layout_upload", "jg_item_upload", "jg_filter_upload", "jg_detect_batch",
                 "jg_collect_batch", "jg_cde_register_hazard", "jg_cde_deregister_hazard", "jg_cde_save", "jg_cde_restore",
                 "jg_cde_flatten"):
        assert must in names


def test_library_exports_every_declared_symbol(jg):
    lib = jg.load_library()
    raw = ctypes.CDLL(jg.lib_path())
    for name in declared_functions():
        assert hasattr(raw, name), f"{name} declared in include/jagua_cuda.h but not exported"
        assert name in jg.C_ABI, f"{name} has no ctypes signature in the host mirror"
    for name in jg.C_ABI:
        assert name in declared_functions(), f"{name} bound in Python but not declared in the header"
    assert lib is jg.load_library()


def _no_gpu():
    try:
        return subprocess_has_no_gpu()
    except Exception:
        return True


def subprocess_has_no_gpu():
    import subprocess
    r = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True)
    return r.returncode != 0 or "GPU" not in r.stdout


@pytest.mark.skipif(not _no_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_without_gpu(jg):
    wl = jg.load_workload("swim_half")
    with pytest.raises(jg.JaguaCudaError):
        jg.build_engine(wl, device=0)
    # host half alone is fine, but a query on it must raise, never silently compute on the CPU
    cde = jg.build_engine(wl, device=None)
    with pytest.raises(jg.JaguaCudaError):
        cde.detect_batch(0, jg.make_queries(wl, 0, 16))
    cde.close()


def test_status_codes_for_bad_arguments(jg):
    lib = jg.load_library()
    cfg = jg.JgConfig(5, 16, 2, 0)
    out = ctypes.c_void_p()
    assert lib.jg_engine_create(0, None, ctypes.byref(out)) == jg.JG_ERR_INVALID
    bad = jg.JgConfig(99, 16, 2, 0)  # deeper than the 15 levels the walk state holds
    assert lib.jg_engine_create(0, ctypes.byref(bad), ctypes.byref(out)) == jg.JG_ERR_INVALID
    import numpy as np
    assert lib.jg_cde_create(np.array([0, 0, -1, 1], np.float32), ctypes.byref(cfg), ctypes.byref(out)) == jg.JG_ERR_INVALID
