"""ctypes binding of libgpb200.so (include/gpb200.h).

There is NO CPU fallback: if the shared library is missing or no CUDA device is visible, the first numerical call raises.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgpb200.so")

KERN_RBF, KERN_MATERN52 = 0, 1
ACQ_EI, ACQ_LCB = 0, 1
ERR_CUDA, ERR_ARG, ERR_DOMAIN = -1, -2, -3
KIND_IDS = {"rbf": KERN_RBF, "RBF": KERN_RBF, "mat52": KERN_MATERN52, "Mat52": KERN_MATERN52, "Matern52": KERN_MATERN52}
ACQ_IDS = {"EI": ACQ_EI, "LCB": ACQ_LCB}

c_double_p = ctypes.POINTER(ctypes.c_double)
c_ll_p = ctypes.POINTER(ctypes.c_longlong)
c_int_p = ctypes.POINTER(ctypes.c_int)
c_void_p = ctypes.c_void_p
c_int = ctypes.c_int

# name -> (restype, argtypes); mirrors include/gpb200.h one to one (tests check that every declared symbol is exported)
SIGNATURES = {
    "gpb_version": (c_int, []),
    "gpb_last_error": (ctypes.c_char_p, []),
    "gpb_device_count": (c_int, []),
    "gpb_launch_count": (ctypes.c_longlong, []),
    "gpb_kern_K": (c_int, [c_int, c_int, c_int, c_void_p, c_int, c_void_p, ctypes.c_double, c_double_p, c_int, c_void_p, c_int,
                           c_int, c_void_p]),
    "gpb_kern_update_gradients_full": (c_int, [c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, ctypes.c_double,
                                               c_double_p, c_int, c_double_p, c_int, c_void_p]),
    "gpb_kern_K_gower": (c_int, [c_int, c_int, c_int, c_void_p, c_int, c_void_p, ctypes.c_double, c_int_p, c_double_p, c_void_p,
                                 c_int, c_int, c_void_p]),
    "gpb_kern_update_gradients_full_gower": (c_int, [c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, ctypes.c_double,
                                                     c_double_p, c_int, c_int_p, c_double_p, c_double_p, c_int, c_void_p]),
    "gpb_model_set_gower": (c_int, [c_void_p, c_int, c_int_p, c_double_p]),
    "gpb_kern_gradients_X": (c_int, [c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int, ctypes.c_double, c_double_p,
                                     c_int, c_void_p, c_int, c_void_p]),
    "gpb_pdinv": (c_int, [c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_double_p, c_int, c_void_p]),
    "gpb_potrs": (c_int, [c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p]),
    "gpb_potri": (c_int, [c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p]),
    "gpb_model_workspace_bytes": (ctypes.c_size_t, [c_int, c_int, c_int, c_int]),
    "gpb_model_create": (c_int, [ctypes.POINTER(c_void_p), c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, ctypes.c_size_t,
                                 c_void_p]),
    "gpb_model_destroy": (c_int, [c_void_p]),
    "gpb_model_set_data": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int]),
    "gpb_model_set_theta": (c_int, [c_void_p, ctypes.c_double, c_double_p, ctypes.c_double]),
    "gpb_model_fit": (c_int, [c_void_p, c_int, ctypes.c_double, c_double_p]),
    "gpb_model_append": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_double_p]),
    "gpb_model_get": (c_int, [c_void_p, ctypes.c_char_p, c_void_p, c_int, c_int]),
    "gpb_model_predict": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int]),
    "gpb_model_predict_full_cov": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int]),
    "gpb_model_predictive_gradients": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int]),
    "gpb_model_fmin": (c_int, [c_void_p, c_double_p]),
    "gpb_model_acquisition": (c_int, [c_void_p, c_int, ctypes.c_double, ctypes.c_double, c_int, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    "gpb_model_set_penalizers": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "gpb_model_acquisition_lp": (c_int, [c_void_p, c_int, ctypes.c_double, ctypes.c_double, c_int, c_void_p, c_void_p, c_void_p,
                                         c_int]),
    "gpb_model_acq_topk": (c_int, [c_void_p, c_int, ctypes.c_double, ctypes.c_double, c_int, c_void_p, c_int, c_int,
                                   ctypes.c_longlong, c_double_p, c_ll_p, c_double_p]),
    "gpb_set_overlap": (c_int, [c_int]),
    "gpb_model_acq_topk_full": (c_int, [c_void_p, c_int, ctypes.c_double, ctypes.c_double, c_int, c_void_p, c_int, c_int,
                                        ctypes.c_longlong, c_double_p, c_ll_p, c_double_p, c_void_p, c_void_p]),
    "gpb_model_acq_topk_dev": (c_int, [c_void_p, c_int, ctypes.c_double, ctypes.c_double, c_int, c_void_p, c_int, ctypes.c_longlong,
                                       c_void_p, c_void_p, c_void_p]),
    "gpb_model_state_ptr": (c_int, [c_void_p, ctypes.c_char_p, ctypes.POINTER(c_void_p), ctypes.POINTER(ctypes.c_size_t)]),
    "gpb_model_adopt_state": (c_int, [c_void_p, ctypes.c_double, c_double_p, ctypes.c_double, ctypes.c_double, c_int]),
    "gpb_profile_gemm": (c_int, [c_int]),
    "gpb_gemm_config": (c_int, [c_int]),
    "gpb_profile_gemm_collect": (c_int, [c_double_p, c_double_p, c_ll_p]),
    "gpb_profile_gemm_last": (c_int, [c_double_p, c_double_p]),
    "gpb_dgemm": (c_int, [c_int, c_int, c_int, c_int, c_int, ctypes.c_double, c_void_p, c_int, c_void_p, c_int, ctypes.c_double,
                          c_void_p, c_int, c_void_p]),
    "gpb_ozaki_dgemm": (c_int, [c_int, c_int, c_int, c_int, c_int, ctypes.c_double, c_void_p, c_int, c_void_p, c_int, ctypes.c_double,
                                c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "gpb_set_ozaki": (c_int, [c_int, c_int]),
    "gpb_model_engine_report": (c_int, [c_void_p, c_int_p, c_double_p]),
    "gpb_ozaki_fallback_count": (ctypes.c_longlong, []),
    "gpb_ozaki_crt_bits": (c_int, [c_int, ctypes.c_longlong]),
    "gpb_ozaki_crt_host_residues": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "gpb_ozaki_crt_host_combine": (c_int, [c_void_p, ctypes.c_longlong, c_int, c_void_p]),
}

_lib = None


class GpbError(RuntimeError):
    pass


def load():
    """Load libgpb200.so (built by `python -c 'import __graft_entry__ as g; g.build()'` or `make -C .../csrc`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GpbError("libgpb200.so not found at %s -- build it (make -C gaussian_process_optimization_b200/csrc); "
                       "this package has no CPU fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error():
    return load().gpb_last_error().decode("utf-8", "replace")


def check(rc, what=""):
    """rc == 0: fine.  rc > 0: LAPACK-style info (not positive definite) -> numpy.linalg.LinAlgError like the reference
    (GPy/GPy/util/linalg.py:64,75).  rc < 0: usage / CUDA error."""
    if rc == 0:
        return
    if rc == ERR_DOMAIN:
        raise np.linalg.LinAlgError("not positive definite, even with jitter. (%s)" % last_error())
    if rc > 0:
        raise np.linalg.LinAlgError("not positive definite, even with jitter." if what == "jitchol" else
                                    "%s: %s" % (what or "gpb", last_error()))
    raise GpbError("%s failed (%d): %s" % (what or "gpb call", rc, last_error()))


def require_gpu():
    lib = load()
    if lib.gpb_device_count() < 1:
        raise GpbError("no CUDA device visible: gaussian_process_optimization_b200 has no CPU fallback")
    return lib


def is_torch(x):
    return type(x).__module__.startswith("torch")


def as_host(x, shape=None):
    a = np.ascontiguousarray(np.asarray(x, dtype=np.float64))
    if shape is not None:
        a = a.reshape(shape)
    return a


def ptr(x):
    """Raw address of a C-contiguous float64 numpy array or CUDA torch tensor (None -> NULL)."""
    if x is None:
        return None
    if is_torch(x):
        assert x.is_contiguous() and str(x.dtype) == "torch.float64" and x.is_cuda
        return ctypes.c_void_p(x.data_ptr())
    assert isinstance(x, np.ndarray) and x.dtype == np.float64 and x.flags["C_CONTIGUOUS"]
    return ctypes.c_void_p(x.ctypes.data)


def dptr(a):
    return a.ctypes.data_as(c_double_p)


def current_stream():
    """Stream handle to launch on: torch's current CUDA stream when torch has initialised CUDA, else the default stream."""
    try:
        import torch
        if torch.cuda.is_available() and torch.cuda.is_initialized():
            return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    except Exception:
        pass
    return ctypes.c_void_p(0)
