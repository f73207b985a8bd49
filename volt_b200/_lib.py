"""ctypes binding of libvolt_b200.so (the C ABI declared in include/volt_b200.h).

The shared library is built in-tree by `__graft_entry__.build()` / `python -m volt_b200.build`.  There is NO CPU
fallback: if the library is missing, or the current device is not an sm_100 GPU, every op raises.
"""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_longlong, c_uint, c_ulonglong, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
# VOLT_B200_LIB: developer override used by tools/ab.sh to time experimental builds of the same ABI side by side
LIB_PATH = os.environ.get("VOLT_B200_LIB") or os.path.join(HERE, "csrc", "libvolt_b200.so")

VOLT_NSCALARS = 16
S_MLL, S_DNOISE, S_LOGDET, S_INVQUAD, S_TRINV, S_ALAL, S_ALR, S_JITTER, S_Z2Z2, S_Z1Z2, S_DRAW, S_NOISE = range(12)
MA_EWMA, MA_DEWMA, MA_TEWMA, MA_MEANREVERT, MA_GIVEN = range(5)
VOL_RAW, VOL_SIGMA, VOL_LOGSIGMA = 0, 1, 2

_fp = c_void_p  # device (or host) pointers travel as integers
_SIGS = {
    "volt_last_error": (c_char_p, []),
    "volt_abi_version": (c_int, []),
    "volt_device_check": (c_int, []),
    "volt_launch_count": (c_longlong, []),
    "volt_set_mll_impl": (c_int, [c_int]),
    "volt_release_workspaces": (c_int, []),
    "volt_cumtrapz": (c_int, [_fp, c_int, _fp, c_int, c_int, c_int, c_int, _fp, c_void_p]),
    "volt_vol_cov": (c_int, [_fp, c_int, _fp, c_int, c_int, c_int, _fp, c_int, _fp, c_void_p]),
    "volt_bm_cov": (c_int, [_fp, c_int, _fp, c_int, _fp, _fp, c_void_p]),
    "volt_ewma": (c_int, [_fp, c_int, c_int, c_int, _fp, c_void_p]),
    "volt_ma_mean": (c_int, [_fp, c_int, c_int, c_int, c_int, c_float, _fp, _fp, _fp, _fp, _fp, c_void_p]),
    "volt_mll_grad_vol": (c_int, [_fp, c_int, _fp, c_int, _fp, _fp, c_int, c_int, c_int, c_float, c_int, _fp, _fp, _fp,
                                  c_void_p]),
    "volt_mll_grad_vol_raw": (c_int, [_fp, c_int, _fp, c_int, _fp, _fp, c_int, c_int, c_int, c_float, c_int, _fp, _fp, _fp, _fp,
                                      c_void_p]),
    "volt_mll_step_sharded": (c_int, [_fp, c_int, _fp, c_int, _fp, _fp, c_int, c_int, c_int, c_float, c_int, _fp, _fp, _fp, _fp,
                                      c_void_p, c_void_p, _fp, c_int, c_int, c_int, c_int, c_uint, c_void_p]),
    "volt_loss_gather": (c_int, [c_void_p, c_int, c_int, c_uint, _fp, c_void_p]),
    "volt_loss_push": (c_int, [_fp, c_void_p, c_int, c_int, c_int, c_uint, c_void_p]),
    "volt_mll_grad_bm": (c_int, [_fp, _fp, c_int, _fp, _fp, c_int, c_int, c_int, c_float, c_int, _fp, _fp, _fp, c_void_p]),
    "volt_mll_grad_bm_inv": (c_int, [_fp, _fp, c_int, _fp, _fp, c_int, c_int, c_int, c_float, c_int, _fp, _fp, _fp, _fp, c_void_p]),
    "volt_mll_grad_dense": (c_int, [_fp, c_longlong, c_int, _fp, _fp, c_int, c_int, c_int, c_float, c_int, _fp, _fp, _fp,
                                    c_void_p]),
    "volt_mll_grad_vol_host": (c_int, [_fp, _fp, _fp, _fp, c_int, c_int, c_int, c_float, c_int, _fp, _fp, _fp]),
    "volt_potrf": (c_int, [_fp, c_longlong, c_int, _fp, c_int, c_int, c_int, c_float, c_int, _fp, c_longlong, c_int, _fp, _fp,
                           c_void_p]),
    "volt_potrs": (c_int, [_fp, c_longlong, c_int, c_int, c_int, _fp, c_longlong, c_int, c_int, c_void_p]),
    "volt_bmgp_posterior": (c_int, [_fp, _fp, c_int, c_int, _fp, c_int, _fp, c_int, _fp, c_int, _fp, _fp, _fp, c_void_p]),
    "volt_mvn_sample": (c_int, [_fp, _fp, _fp, c_int, c_int, c_int, c_float, c_int, _fp, _fp, c_void_p]),
    "volt_rollout": (c_int, [_fp, _fp, _fp, c_int, _fp, _fp, c_int, c_int, c_int, c_int, c_int, c_int, c_float, _fp, _fp, _fp,
                             c_int, c_float, _fp, c_int, c_float, c_ulonglong, _fp, _fp, _fp, c_void_p]),
    "volt_rollout_normals": (c_int, [c_ulonglong, c_int, c_int, c_int, c_int, _fp, c_void_p]),
    "volt_rollout_stats": (c_int, [_fp, c_int, c_int, c_int, _fp, _fp, c_int, _fp, _fp, _fp, _fp, _fp, c_void_p]),
    "volt_gpcv_rows": (c_int, [_fp, _fp, _fp, _fp, _fp, _fp, c_int, c_int, c_int, c_float, _fp, _fp, c_void_p]),
    "volt_gemm_nt": (c_int, [_fp, c_longlong, c_longlong, _fp, c_longlong, c_longlong, _fp, c_longlong, c_longlong, c_int, c_int,
                             c_int, c_int, c_int, c_void_p]),
    "volt_adam_step": (c_int, [_fp, _fp, _fp, _fp, c_longlong, c_float, c_float, c_float, c_float, c_int, _fp, c_void_p]),
}

EXPORTED = tuple(_SIGS)
_lib = None


class VoltLibraryError(RuntimeError):
    pass


def load():
    """Load libvolt_b200.so once and attach prototypes.  Raises VoltLibraryError when the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VoltLibraryError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(volt_b200 has no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.volt_abi_version() != 1:
        raise VoltLibraryError("libvolt_b200.so ABI version mismatch")
    _lib = lib
    return lib


def last_error():
    return load().volt_last_error().decode("utf-8", "replace")


def check(status, what):
    if status != 0:
        raise VoltLibraryError(f"{what} failed with status {status}: {last_error()}")


def require_device():
    """Fail loudly unless torch sees a CUDA device and the library accepts it (sm_100)."""
    import torch

    if not torch.cuda.is_available():
        raise VoltLibraryError("volt_b200 needs a CUDA device (B200, sm_100); there is no CPU fallback")
    check(load().volt_device_check(), "volt_device_check")


def launch_count():
    return int(load().volt_launch_count())
