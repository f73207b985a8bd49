"""torch.ops.volt.* -- the TORCH_LIBRARY registration of the hot-path operators ("Extension ops").

`load()` loads volt_b200/csrc/libvolt_torch.so (built by __graft_entry__.build() from csrc/torch_shim.cpp), a thin shim
over the same plain-C ABI the ctypes binding uses (libvolt_b200.so).  After it:

    torch.ops.volt.vol_cov(x, vol, add_diag=None)                  -> K (B,T,T)
    torch.ops.volt.bm_cov(x1, x2, vol)                             -> K (n1,n2)
    torch.ops.volt.ewma(y, k, mode=0)                              -> (.., T+1)
    torch.ops.volt.potrf_(A)                                       -> info (B)        (A overwritten with its factor)
    torch.ops.volt.mll_fwd_bwd(x, vol, resid, noise)               -> (mll, dnoise, alpha, logdet)
    torch.ops.volt.gp_predict(L, Kx, r)                            -> (mean, cov_reduction)
    torch.ops.volt.rollout(x_train, y_train, vol_train, test_x, pred_vol, eps=None, k=25, theta=None, latent_mean=None, seed=0)

All inputs are CUDA tensors; there is no CPU implementation registered (calling with CPU tensors raises).
"""
import os

import torch

from . import _lib

SHIM_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libvolt_torch.so")
_loaded = False


def load():
    """Register the `volt` operator namespace (idempotent).  Raises VoltLibraryError when the shim is not built."""
    global _loaded
    if _loaded:
        return torch.ops.volt
    _lib.load()          # libvolt_b200.so first: the shim links against it
    if not os.path.exists(SHIM_PATH):
        raise _lib.VoltLibraryError(f"{SHIM_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    torch.ops.load_library(SHIM_PATH)
    _loaded = True
    return torch.ops.volt
