"""ctypes binding of include/femflow_mpm.h.  No fallback: if the CUDA library is
missing the import of this module fails loudly."""
from __future__ import annotations

import ctypes as C
import os

from ._build import LIBPATH

FFMPM_OK = 0
FFMPM_E_INVALID, FFMPM_E_CUDA, FFMPM_E_OOB, FFMPM_E_STATE = -1, -2, -3, -4
FFMPM_F32, FFMPM_F64 = 0, 1
FFMPM_NEO_HOOKEAN, FFMPM_SNOW = 0, 1
FFMPM_P2G_AUTO, FFMPM_P2G_SCATTER, FFMPM_P2G_TILED, FFMPM_P2G_FUSED = 0, 1, 2, 3
ABI_VERSION = 5


class FfMpmConfig(C.Structure):
    _fields_ = [
        ("dim", C.c_int32), ("dtype", C.c_int32), ("model", C.c_int32),
        ("res", C.c_int32 * 3), ("n", C.c_int32 * 3), ("origin", C.c_int32 * 3),
        ("inv_dx", C.c_double), ("dx", C.c_double), ("dt", C.c_double), ("volume", C.c_double),
        ("gravity", C.c_double), ("hardening", C.c_double),
        ("mass", C.c_double), ("mu_0", C.c_double), ("lambda_0", C.c_double),
        ("p2g_mode", C.c_int32), ("reserved", C.c_int32 * 7),
    ]


class FfMpmState(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("v", C.c_void_p), ("C", C.c_void_p), ("F", C.c_void_p), ("Jp", C.c_void_p),
        ("mass", C.c_void_p), ("mu0", C.c_void_p), ("lam0", C.c_void_p), ("id", C.c_void_p),
        ("material", C.c_void_p), ("stride", C.c_int64),
    ]


# every symbol include/femflow_mpm.h declares: name -> (restype, argtypes)
H = C.c_void_p
PROTOTYPES = {
    "ffmpm_abi_version": (C.c_int32, []),
    "ffmpm_last_error": (C.c_char_p, []),
    "ffmpm_workspace_bytes": (C.c_int64, [C.POINTER(FfMpmConfig), C.c_int64]),
    "ffmpm_create": (C.c_int, [C.POINTER(FfMpmConfig), C.c_int32, C.POINTER(H)]),
    "ffmpm_destroy": (None, [H]),
    "ffmpm_set_workspace": (C.c_int, [H, C.c_void_p, C.c_int64]),
    "ffmpm_bind_state": (C.c_int, [H, C.POINTER(FfMpmState), C.POINTER(FfMpmState), C.c_int64]),
    "ffmpm_live_buffer": (C.c_int, [H]),
    "ffmpm_num_particles": (C.c_int64, [H]),
    "ffmpm_set_num_particles": (C.c_int, [H, C.c_int64]),
    "ffmpm_clear_grid": (C.c_int, [H, C.c_void_p]),
    "ffmpm_bin": (C.c_int, [H, C.c_void_p]),
    "ffmpm_p2g": (C.c_int, [H, C.c_void_p]),
    "ffmpm_grid_op": (C.c_int, [H, C.c_void_p]),
    "ffmpm_g2p": (C.c_int, [H, C.c_void_p]),
    "ffmpm_grid_op_halo": (C.c_int, [H, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p]),
    "ffmpm_scatter": (C.c_int, [H, C.c_void_p]),
    "ffmpm_gather": (C.c_int, [H, C.c_void_p]),
    "ffmpm_substep": (C.c_int, [H, C.c_int32, C.c_void_p]),
    "ffmpm_set_materials": (C.c_int, [H, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int32]),
    "ffmpm_set_colliders": (C.c_int, [H, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_int32]),
    "ffmpm_collide": (C.c_int, [H, C.c_void_p]),
    "ffmpm_set_owned_range": (C.c_int, [H, C.c_int32, C.c_int32]),
    "ffmpm_set_owned_slack": (C.c_int, [H, C.c_int32]),
    "ffmpm_leaver_count_ptr": (C.c_int, [H, C.POINTER(C.c_void_p)]),
    "ffmpm_migrate_rows": (C.c_int32, [H]),
    "ffmpm_migrate_pack": (C.c_int, [H, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "ffmpm_migrate_unpack": (C.c_int, [H, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "ffmpm_grid_ptr": (C.c_int, [H, C.POINTER(C.c_void_p)]),
    "ffmpm_grid_view": (C.c_int, [H, C.POINTER(C.c_void_p)]),
    "ffmpm_bin_ptrs": (C.c_int, [H, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                 C.POINTER(C.c_int64)]),
    "ffmpm_poll_error": (C.c_int, [H, C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int64)]),
    "ffmpm_snapshot": (C.c_int, [H, C.c_double, C.c_void_p, C.c_void_p]),
    "ffmpm_export_state": (C.c_int, [H, C.POINTER(FfMpmState), C.c_void_p]),
    "ffmpm_scene_scratch_bytes": (C.c_int64, [C.c_int32]),
    "ffmpm_gen_implicit_points": (C.c_int, [C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "ffmpm_gen_cube_points": (C.c_int, [C.POINTER(C.c_double), C.c_int32, C.c_void_p, C.c_void_p]),
    "ffmpm_launch_count": (C.c_int64, [H]),
    "ffmpm_debug_red_add4": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_int64, C.c_void_p]),
}


def load_library(path: str = "") -> C.CDLL:
    # FEMFLOW_MPM_LIB: load another build of the same ABI (A/B runs of kernel variants on one box)
    override = os.environ.get("FEMFLOW_MPM_LIB")
    if not path and override:
        import sys
        print(f"femflow_b200: loading the CUDA library from FEMFLOW_MPM_LIB={override} instead of {LIBPATH}", file=sys.stderr)
    path = path or override or LIBPATH
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: the femflow_b200 CUDA library has not been built. "
            "Run `python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
            "There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    got = lib.ffmpm_abi_version()
    if got != ABI_VERSION:
        raise ImportError(f"libfemflow_mpm.so ABI {got} != binding ABI {ABI_VERSION}; rebuild")
    return lib


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = load_library()
    return _lib


class MpmError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"femflow_mpm error {code}: {msg}")
        self.code = code


def check(rc: int) -> None:
    if rc != FFMPM_OK:
        msg = lib().ffmpm_last_error().decode(errors="replace")
        raise MpmError(rc, msg)
