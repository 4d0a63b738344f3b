"""GPy.util.linalg surface used around exact inference (GPy/GPy/util/linalg.py), backed by libgpb200.so.

jitchol (:56-81), pdinv (:193-214), dpotrs (:116-125), dpotri (:127-145), dtrtri (:217-227), symmetrify (:356-379).
The jitter ladder lives here, on the host, exactly as in the reference: first jitter = mean(diag) * 1e-6, x10 per retry, at
most `maxtries` retries, LinAlgError otherwise (and immediately when a diagonal entry is not positive).
"""
import numpy as np

from . import native


def _with_jitter(A, maxtries, want):
    A = np.ascontiguousarray(A, dtype=np.float64)
    rc, Ai, L, Li, logdet = native.pdinv(A, want=want)
    if rc == 0:
        return Ai, L, Li, logdet
    if rc < 0:
        native._lib.check(rc, "pdinv")
    diagA = np.diag(A)
    if np.any(diagA <= 0.):
        raise np.linalg.LinAlgError("not pd: non-positive diagonal elements")
    jitter = diagA.mean() * 1e-6
    num_tries = 1
    while num_tries <= maxtries and np.isfinite(jitter):
        rc, Ai, L, Li, logdet = native.pdinv(A + np.eye(A.shape[0]) * jitter, want=want)
        if rc == 0:
            return Ai, L, Li, logdet
        jitter *= 10
        num_tries += 1
    raise np.linalg.LinAlgError("not positive definite, even with jitter.")


def jitchol(A, maxtries=5):
    return _with_jitter(A, maxtries, ("L",))[1]


def pdinv(A, *args):
    """-> (Ai, L, Li, logdet)"""
    maxtries = args[0] if args else 5
    return _with_jitter(A, maxtries, ("Ai", "L", "Li", "logdet"))


def dpotrs(A, B, lower=1):
    assert lower == 1, "only lower factors are produced on this path"
    return native.potrs(A, B), 0


def dpotri(A, lower=1):
    assert lower == 1, "only lower factors are produced on this path"
    return native.potri(A), 0


def dtrtri(L):
    raise NotImplementedError("dtrtri's result is unused by exact inference (linalg.py:209 vs exact_gaussian_inference.py:58); "
                              "use pdinv(A)[2] for L^-1")


def symmetrify(A, upper=False):
    """In place, like linalg_cython.pyx:9-18 (a host-side index copy: no arithmetic)."""
    if not upper:
        iu = np.triu_indices_from(A, k=1)
        A[iu] = A.T[iu]
    else:
        il = np.tril_indices_from(A, k=-1)
        A[il] = A.T[il]
