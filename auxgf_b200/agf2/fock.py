"""Inner self-consistency of the Fock matrix (auxgf/agf2/fock.py:41-149 RHF, 152-276 UHF): alternate the
chemical-potential optimisation with density/Fock updates until the density and the electron number are
converged.  Mutates ``se`` and ``rdm`` in place like the reference (the drivers ignore the return values,
optragf2.py:301)."""
import numpy as np

from auxgf_b200.util import DIIS
from auxgf_b200.agf2.chempot import minimize, diag_fock_ext


def fock_loop_rhf(se, hf, rdm, fock_func=None, diis_space=8, netol=1e-6, dtol=1e-6, verbose=False, maxiter=50,
                  maxruns=20, jac=True, log=None):
    get_fock = fock_func if fock_func is not None else (lambda r: hf.get_fock(r, basis="mo"))
    diis = DIIS(space=diis_space)
    fock = get_fock(rdm)
    nphys = se.nphys
    buf = np.zeros((nphys + se.naux,) * 2)
    rmsd, error = np.inf, np.inf
    for nrun in range(1, maxruns + 1):
        se, opt = minimize(se, fock, hf.nelec, buf=buf, x0=se.chempot, tol=netol, jac=jac)
        rdm_prev = None
        for niter in range(1, maxiter + 1):
            w, v, se.chempot, error = diag_fock_ext(se, fock, hf.nelec)
            c_occ = v[:nphys, w < se.chempot]
            rdm[:, :] = (c_occ @ c_occ.T) * 2
            fock = get_fock(rdm)
            if niter > 1:
                fock = diis.update(fock)
                rmsd = np.sqrt(np.sum((rdm - rdm_prev) ** 2))
                if rmsd < dtol:
                    break
            rdm_prev = rdm.copy()
        if log is not None:
            log("%6d %12.6g %6d %12.6g %12.6f" % (nrun, rmsd, niter, error, se.chempot))
        if rmsd < dtol and abs(error) < netol:
            break
    return se, rdm, bool(rmsd < dtol and abs(error) < netol)


def fock_loop_uhf(se, hf, rdm, fock_func=None, diis_space=8, netol=1e-6, dtol=1e-6, verbose=False, maxiter=50,
                  maxruns=20, jac=True, log=None):
    """Spin channels are optimised separately (occupancy 1, electron numbers nalph / nbeta)."""
    get_fock = fock_func if fock_func is not None else (lambda r: hf.get_fock(r, basis="mo"))
    diis = DIIS(space=diis_space)
    fock = get_fock(rdm)
    nphys = se[0].nphys
    nel = (hf.nalph, hf.nbeta)
    bufs = [np.zeros((nphys + s.naux,) * 2) for s in se]
    rmsd, errors = np.inf, [np.inf, np.inf]
    for nrun in range(1, maxruns + 1):
        for s in range(2):
            minimize(se[s], fock[s], nel[s], occupancy=1.0, buf=bufs[s], x0=se[s].chempot, tol=netol, jac=jac)
        rdm_prev = None
        for niter in range(1, maxiter + 1):
            for s in range(2):
                w, v, se[s].chempot, errors[s] = diag_fock_ext(se[s], fock[s], nel[s], occupancy=1.0)
                c_occ = v[:nphys, w < se[s].chempot]
                rdm[s][:, :] = c_occ @ c_occ.T
            fock = get_fock(rdm)
            if niter > 1:
                fock = diis.update(fock)
                rmsd = np.sqrt(np.sum((rdm - rdm_prev) ** 2))
                if rmsd < dtol:
                    break
            rdm_prev = rdm.copy()
        if log is not None:
            log("%6d %12.6g %6d %12.6g %12.6g" % (nrun, rmsd, niter, errors[0], errors[1]))
        if rmsd < dtol and max(abs(e) for e in errors) < netol:
            break
    return se, rdm, bool(rmsd < dtol and max(abs(e) for e in errors) < netol)
