"""Chemical-potential optimisation of the self-energy poles (auxgf/agf2/chempot.py:11-225).

The shift ``x`` applied to the auxiliary energies is chosen so that the Aufbau filling of the extended
Fock matrix holds ``nelec`` physical electrons; objective = (electron-number error)^2 with the analytic
gradient of chempot.py:77-118, minimised by SciPy TNC.  Every objective/gradient evaluation is one dense
eigensolve of the (nphys+naux)^2 extended Fock matrix, on the GPU through ``Aux.eig``."""
import numpy as np
from scipy import optimize

from auxgf_b200.util import find_chempot


def diag_fock_ext(se, fock, nelec, chempot=0.0, occupancy=2.0, buf=None):
    """-> (w, v_phys, chempot, error): eigenvalues and the PHYSICAL rows of the eigenvectors of the extended Fock
    matrix (the auxiliary rows never leave the GPU), Aufbau chemical potential and electron-number error."""
    w, v = se.eig_phys(fock, chempot=chempot, out=buf)
    chempot_phys, error = find_chempot(se.nphys, nelec, (w, v), occupancy=occupancy)
    return w, v, chempot_phys, error


class _LastSolve:
    """SciPy asks for the objective and for its gradient at the same point in two separate calls (so does the
    reference, chempot.py:187-190: two eigensolves per point); remember the last eigensolve and reuse it."""

    def __init__(self):
        self.x = None
        self.result = None

    def __call__(self, x, se, fock, nelec, occupancy, buf):
        if self.x != x:
            self.result = diag_fock_ext(se, fock, nelec, chempot=x, occupancy=occupancy, buf=buf)
            self.x = x
        return self.result


def objective(x, se, fock, nelec, occupancy=2.0, buf=None, memo=None):
    x = float(np.ravel(x)[0])
    solve = memo if memo is not None else _LastSolve()
    return solve(x, se, fock, nelec, occupancy, buf)[3] ** 2


def gradient(x, se, fock, nelec, occupancy=2.0, buf=None, return_val=False, memo=None):
    x = float(np.ravel(x)[0])
    solve = memo if memo is not None else _LastSolve()
    w, v, chempot, error = solve(x, se, fock, nelec, occupancy, buf)
    nocc = int(np.sum(w < chempot))
    n = se.nphys
    # chempot.py:96-99 uses the auxiliary rows:  h_1 = -v_aux[:, vir]^T v_aux[:, occ].  The eigenvectors are
    # orthonormal, so for an occupied / virtual pair  v_aux^T v_aux = -v_phys^T v_phys : the physical rows suffice
    h_1 = v[:n, nocc:].T @ v[:n, :nocc]
    z_ai = -h_1 / (w[None, :nocc] - w[nocc:, None])
    c_occ = v[:n, nocc:] @ z_ai
    rdm1 = (v[:n, :nocc] @ c_occ.T) * 4
    d = 2 * error * np.trace(rdm1)
    return (error ** 2, d) if return_val else d


def minimize(se, fock, nelec, occupancy=2.0, buf=None, x0=0.0, tol=1e-6, maxiter=200, jac=True):
    """Shifts ``se`` in place (energies -= x, chempot updated) and returns (se, opt)."""
    tol2 = tol ** 2
    memo = _LastSolve()
    args = (se, fock, nelec, occupancy, buf, memo)
    kwargs = dict(x0=np.array([x0]), method="TNC", bounds=[(None, None)],
                  options=dict(maxfun=maxiter, ftol=tol2, xtol=tol2, gtol=tol2))
    if jac:
        kwargs["jac"] = lambda x, *a: np.array([gradient(x, *a[:5], memo=a[5])])
    opt = optimize.minimize(objective, args=args, **kwargs)
    se._ener -= float(np.ravel(opt.x)[0])
    se.chempot = find_chempot(se.nphys, nelec, se.eig_phys(fock), occupancy=occupancy)[0]
    return se, opt
