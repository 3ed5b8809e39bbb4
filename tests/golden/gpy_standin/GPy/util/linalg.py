"""GPy.util.linalg: the four routines exact inference uses (published behaviour of GPy 1.10)."""
import numpy as np
from scipy import linalg as sla
from scipy.linalg import lapack


def tdot(X):
    return np.dot(X, X.T)


def jitchol(A, maxtries=5):
    """dpotrf; on failure retry with mean(diag)*1e-6 * 10^t on the diagonal, t = 0..maxtries-1."""
    A = np.ascontiguousarray(A)
    L, info = lapack.dpotrf(A, lower=1)
    if info == 0:
        jitchol.last_tries = 0
        return L
    diagA = np.diag(A)
    if np.any(diagA <= 0.):
        raise sla.LinAlgError("not pd: non-positive diagonal elements")
    jitter = diagA.mean() * 1e-6
    num_tries = 1
    while num_tries <= maxtries and np.isfinite(jitter):
        try:
            L = sla.cholesky(A + np.eye(A.shape[0]) * jitter, lower=True)
            jitchol.last_tries = num_tries
            return L
        except sla.LinAlgError:
            jitter *= 10
        finally:
            num_tries += 1
    raise sla.LinAlgError("not positive definite, even with jitter.")


jitchol.last_tries = 0


def dtrtrs(A, B, lower=1, trans=0, unitdiag=0):
    return lapack.dtrtrs(np.asfortranarray(A), B, lower=lower, trans=trans, unitdiag=unitdiag)


def dpotrs(A, B, lower=1):
    return lapack.dpotrs(np.asfortranarray(A), B, lower=lower)


def dpotri(A, lower=1):
    R, info = lapack.dpotri(np.asfortranarray(A), lower=lower)
    # symmetrify the triangle LAPACK filled
    tri = np.tril(R) if lower else np.triu(R)
    return tri + tri.T - np.diag(np.diag(tri)), info


def pdinv(A):
    """Returns (A^-1, L, L^-1, log|A|) like GPy."""
    L = jitchol(A)
    logdet = 2. * np.sum(np.log(np.diag(L)))
    Li = dtrtrs(L, np.eye(L.shape[0]), lower=1)[0]
    Ai, _ = dpotri(L, lower=1)
    return Ai, L, Li, logdet
