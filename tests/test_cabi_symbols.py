"""The C-ABI library builds for sm_100a, loads without a GPU and exports every
symbol include/femflow_mpm.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "femflow_mpm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ffmpm_[a-z_0-9]+)\s*\(", text)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for need in ("ffmpm_create", "ffmpm_bind_state", "ffmpm_substep", "ffmpm_bin", "ffmpm_p2g", "ffmpm_grid_op",
                 "ffmpm_g2p", "ffmpm_poll_error", "ffmpm_snapshot", "ffmpm_destroy"):
        assert need in syms


def test_library_exports_every_declared_symbol():
    from femflow_b200._build import LIBPATH, build_library
    build_library()
    lib = ctypes.CDLL(LIBPATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in femflow_mpm.h but not exported"
    from femflow_b200 import _native
    assert set(_native.PROTOTYPES) == set(declared_symbols())
    lib.ffmpm_abi_version.restype = ctypes.c_int32
    assert lib.ffmpm_abi_version() == _native.ABI_VERSION


def test_config_validation_without_gpu():
    """Host-side argument checking works without a device."""
    from femflow_b200 import _native as N
    lib = N.lib()
    cfg = N.FfMpmConfig()
    cfg.dim = 3
    for i in range(3):
        cfg.res[i] = 64; cfg.n[i] = 65
    cfg.dx, cfg.inv_dx, cfg.dt = 1 / 64, 64.0, 1e-4
    b1 = lib.ffmpm_workspace_bytes(ctypes.byref(cfg), 1000)
    b2 = lib.ffmpm_workspace_bytes(ctypes.byref(cfg), 2000)
    assert b1 > 65 ** 3 * 16 and b2 > b1
    cfg.dim = 4
    assert lib.ffmpm_workspace_bytes(ctypes.byref(cfg), 1000) == N.FFMPM_E_INVALID
    assert b"dim" in lib.ffmpm_last_error()
    cfg.dim = 3
    cfg.dt = 0.0
    assert lib.ffmpm_workspace_bytes(ctypes.byref(cfg), 1000) == N.FFMPM_E_INVALID


def test_product_path_fails_loudly_without_gpu():
    torch = pytest.importorskip("torch")
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from femflow_b200.mpm import MpmSolver
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        MpmSolver(3, 32, 1e-4, 1.0, -9.8, 1.0, capacity=10)


def test_no_oracle_import_in_product_package():
    """The product package must never route through oracle/ (checker only)."""
    pkg = os.path.join(ROOT, "femflow_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("# oracle", ""), os.path.join(dirpath, f)
