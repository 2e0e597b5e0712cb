"""ctypes binding of librmhmc_b200.so (include/rmhmc_b200.h).

There is no CPU fallback: if the library is missing or no B200 is visible, every compute entry
point raises.  ``load()`` only dlopens the library (works on a CPU-only box, which is how the
``-m "not gpu"`` tests check the exported symbols); handles are only created on a GPU.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_uint64, c_void_p

from . import build as _build

_LIB = None

# name -> (restype, argtypes); must list every symbol declared in include/rmhmc_b200.h
SIGNATURES = {
    "rmhmc_version": (c_char_p, []),
    "rmhmc_create": (c_int, [POINTER(c_void_p), c_int, c_int64, c_int, c_double, c_void_p, c_void_p]),
    "rmhmc_destroy": (None, [c_void_p]),
    "rmhmc_last_error": (c_char_p, [c_void_p]),
    "rmhmc_update_data": (c_int, [c_void_p, c_void_p, c_void_p]),
    "rmhmc_set_partials_mode": (c_int, [c_void_p, c_int]),
    "rmhmc_get_partials_mode": (c_int, [c_void_p]),
    "rmhmc_set_metric_mode": (c_int, [c_void_p, c_int]),
    "rmhmc_get_metric_mode": (c_int, [c_void_p]),
    "rmhmc_set_launch_regime": (c_int, [c_void_p, c_int]),
    "rmhmc_comm_unique_id": (c_int, [c_char_p]),
    "rmhmc_comm_init": (c_int, [c_void_p, c_int, c_int, c_char_p]),
    "rmhmc_set_stream": (c_int, [c_void_p, c_void_p]),
    "rmhmc_stats_comm_init": (c_int, [c_void_p, c_int, c_int, c_char_p]),
    "rmhmc_stats_gather": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p,
                                   c_void_p, c_int]),
    "blr_device_peaks": (c_int, [c_int, c_void_p, POINTER(c_double), POINTER(c_double)]),
    "rmhmc_metric": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rmhmc_metric_partials": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "rmhmc_chol_logdet": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rmhmc_leapfrog": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_int,
                               c_void_p, c_void_p, c_void_p, c_void_p]),
    "rmhmc_chains_init": (c_int, [c_void_p, c_int64, c_void_p]),
    "rmhmc_configure": (c_int, [c_void_p, c_int, c_double, c_int]),
    "rmhmc_set_tape": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rmhmc_set_philox": (c_int, [c_void_p, c_uint64, c_int64]),
    "rmhmc_set_momentum_family": (c_int, [c_void_p, c_int]),
    "rmhmc_set_tape_chi": (c_int, [c_void_p, c_void_p]),
    "rmhmc_set_samples": (c_int, [c_void_p, c_void_p, c_int64, c_int64]),
    "rmhmc_set_trace": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rmhmc_advance": (c_int, [c_void_p, c_int64, c_int64]),
    "rmhmc_run": (c_int, [c_void_p, c_int64, POINTER(c_int64)]),
    "rmhmc_read_state": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "rmhmc_launch_count": (c_int64, [c_void_p]),
    "rmhmc_chain_generation": (c_int64, [c_void_p]),
    "rmhmc_profile_enable": (c_int, [c_void_p, c_int]),
    "rmhmc_profile_read": (c_int, [c_void_p, c_int, POINTER(c_double), POINTER(c_int64)]),
    "hmc_chains_init": (c_int, [c_void_p, c_int64, c_void_p]),
    "hmc_configure": (c_int, [c_void_p, c_int, c_double]),
    "hmc_set_tape": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "hmc_run": (c_int, [c_void_p, c_int64, POINTER(c_int64)]),
    "hmc_advance": (c_int, [c_void_p, c_int64, c_int64]),
    "hmc_set_fused": (c_int, [c_void_p, c_int, c_int]),
    "mmala_chains_init": (c_int, [c_void_p, c_int64, c_void_p, c_int, c_double]),
    "mmala_set_tape": (c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p]),
    "mmala_run": (c_int, [c_void_p, c_int64, POINTER(c_int64)]),
    "mmala_advance": (c_int, [c_void_p, c_int64, c_int64]),
    "mmala_read_proposal": (c_int, [c_void_p, c_void_p, c_void_p]),
    "blr_ess_batched": (c_int, [c_int, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int64, c_int64, c_int64, c_void_p]),
    "blr_autocorr": (c_int, [c_int, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_void_p]),
    "blr_rhat": (c_int, [c_int, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int64, c_int64, c_void_p]),
    "blr_ess_ragged": (c_int, [c_int, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
}


class RmhmcError(RuntimeError):
    pass


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """dlopen librmhmc_b200.so and attach the signatures; raises if it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.isfile(path):
        raise RmhmcError(
            f"{path} is missing: build it with `python -m riemannhamiltonianmontecarlo_b200.build` "
            "(needs nvcc); this package has no CPU fallback")
    lib = ctypes.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header / library mismatch
        fn.restype = restype
        fn.argtypes = argtypes
    _LIB = lib
    return lib


def check(rc: int, handle=None, what: str = ""):
    if rc == 0:
        return
    lib = load()
    msg = lib.rmhmc_last_error(handle)
    raise RmhmcError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")


def require_cuda():
    import torch

    if not torch.cuda.is_available():
        raise RmhmcError("no CUDA device visible: the RMHMC/HMC samplers only run on a B200 (sm_100a); "
                         "there is no CPU fallback")
    return torch
