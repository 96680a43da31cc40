"""Host-side input providers honouring the ``hf`` protocol the AGF2 drivers consume
(auxgf/util/method.py:74-105, auxgf/agf2/optragf2.py:157-170,325-326):

    h1e_mo, eri_mo, rdm1_mo, e, occ, chempot, nao, nelec, nalph, nbeta, e_tot, e_nuc, with_df,
    mol.{e_nuc, basis, natom, labels, coords}, get_fock(rdm1, basis='mo'), energy_1body(h1e, rdm1, fock)

PySCF is the real provider (``from_pyscf``: any converged density-fitted ``pyscf.scf.RHF/UHF``); it is
not installable in this image, so ``ModelRHF``/``ModelUHF`` build a seeded *model molecule* instead: a
symmetric DF tensor L[Q,p,q] with a decaying spectrum, a diagonal-dominant core Hamiltonian and a
NumPy DF-SCF run to self-consistency, which yields mutually consistent MO energies, integrals and
electron count.  Everything downstream only sees arrays."""
import numpy as np


class Molecule:
    def __init__(self, e_nuc=0.0, basis="model", natom=0):
        self.e_nuc = float(e_nuc)
        self.basis = basis
        self.natom = natom
        self.labels = []
        self.coords = np.zeros((0, 3))


def _model_integrals(nao, naux, seed, strength):
    rng = np.random.default_rng(seed)
    g = rng.standard_normal((naux, nao, nao))
    decay = np.exp(-4.0 * np.arange(naux) / max(naux, 1))
    lpq = strength * decay[:, None, None] * 0.5 * (g + g.transpose(0, 2, 1)) / np.sqrt(nao)
    lpq[0] += 0.6 * np.eye(nao)               # a totally symmetric "charge" component
    h = rng.standard_normal((nao, nao)) * 0.05
    h1e = 0.5 * (h + h.T) + np.diag(np.linspace(-3.0, 2.0, nao))
    return lpq, h1e


def _jk(lpq, dm):
    """J = sum_Q L_Q tr(L_Q D), K = sum_Q L_Q D L_Q as BLAS calls (L_Q symmetric)."""
    naux, nao = lpq.shape[0], lpq.shape[1]
    l2 = lpq.reshape(naux, nao * nao)
    j = (l2.T @ (l2 @ dm.reshape(-1))).reshape(nao, nao)
    tmp = np.matmul(lpq, dm)                                   # (Q, p, s) = L_Q D
    k = tmp.transpose(1, 0, 2).reshape(nao, naux * nao) @ lpq.transpose(1, 0, 2).reshape(nao, naux * nao).T
    return j, k


class ModelRHF:
    """Closed-shell model molecule: ``ModelRHF(nao, nocc, naux, seed).run()``."""
    with_df = True

    def __init__(self, nao=12, nocc=4, naux=30, seed=0, strength=0.5, e_nuc=1.0):
        self.nao, self.nocc_hf, self.naux = nao, nocc, naux
        self.seed, self.strength = seed, strength
        self.mol = Molecule(e_nuc=e_nuc)
        self._done = False

    def run(self, tol=1e-12, maxiter=500):
        lpq, h1e = _model_integrals(self.nao, self.naux, self.seed, self.strength)
        nocc = self.nocc_hf
        e, c = np.linalg.eigh(h1e)
        dm = 2 * c[:, :nocc] @ c[:, :nocc].T
        for it in range(maxiter):
            j, k = _jk(lpq, dm)
            f = h1e + j - 0.5 * k
            e, c = np.linalg.eigh(f)
            dm_new = 2 * c[:, :nocc] @ c[:, :nocc].T
            if np.abs(dm_new - dm).max() < tol:
                dm = dm_new
                break
            dm = 0.5 * dm + 0.5 * dm_new if it < 5 else dm_new
        j, k = _jk(lpq, dm)
        f = h1e + j - 0.5 * k
        self.e, self.c = np.linalg.eigh(f)
        self.occ = np.zeros(self.nao); self.occ[:nocc] = 2.0
        self._lpq_ao, self._h1e_ao = lpq, h1e
        self.e_nuc = self.mol.e_nuc
        self.e_elec = 0.5 * np.sum(dm * (h1e + f))
        self.e_tot = self.e_elec + self.e_nuc
        self._done = True
        return self

    # ---- protocol ------------------------------------------------------------------------------
    @property
    def nelec(self):
        return 2 * self.nocc_hf

    @property
    def nalph(self):
        return self.nocc_hf

    @property
    def nbeta(self):
        return self.nocc_hf

    @property
    def chempot(self):
        return 0.5 * (self.e[self.nocc_hf - 1] + self.e[self.nocc_hf])

    @property
    def h1e_mo(self):
        return self.c.T @ self._h1e_ao @ self.c

    @property
    def eri_mo(self):
        """(naux, nao, nao) DF tensor in the MO basis (hf/hf.py:238-242)."""
        return np.einsum("Qpq,pi,qj->Qij", self._lpq_ao, self.c, self.c, optimize=True)

    @property
    def rdm1_mo(self):
        return np.diag(self.occ)

    def get_fock(self, rdm1, basis="mo"):
        cc = self.c if basis == "mo" else np.eye(self.nao)
        dm = cc @ rdm1 @ cc.T
        j, k = _jk(self._lpq_ao, dm)
        f = self._h1e_ao + j - 0.5 * k
        return cc.T @ f @ cc

    @staticmethod
    def energy_1body(h1e, rdm1, fock):
        return 0.5 * np.sum(rdm1 * (h1e + fock))


class ModelUHF:
    """Open-shell model molecule (nalph != nbeta allowed)."""
    with_df = True

    def __init__(self, nao=12, nalph=4, nbeta=3, naux=30, seed=0, strength=0.5, e_nuc=1.0):
        self.nao, self.nalph, self.nbeta, self.naux = nao, nalph, nbeta, naux
        self.seed, self.strength = seed, strength
        self.mol = Molecule(e_nuc=e_nuc)

    def run(self, tol=1e-12, maxiter=1000):
        lpq, h1e = _model_integrals(self.nao, self.naux, self.seed, self.strength)
        nel = (self.nalph, self.nbeta)
        e0, c0 = np.linalg.eigh(h1e)
        dm = np.stack([c0[:, :n] @ c0[:, :n].T for n in nel])
        for it in range(maxiter):
            f = self._fock_ao(lpq, h1e, dm)
            ec = [np.linalg.eigh(f[s]) for s in range(2)]
            dm_new = np.stack([ec[s][1][:, :nel[s]] @ ec[s][1][:, :nel[s]].T for s in range(2)])
            if np.abs(dm_new - dm).max() < tol:
                dm = dm_new
                break
            dm = 0.5 * dm + 0.5 * dm_new if it < 10 else dm_new
        f = self._fock_ao(lpq, h1e, dm)
        ec = [np.linalg.eigh(f[s]) for s in range(2)]
        self.e = np.stack([ec[0][0], ec[1][0]])
        self.c = np.stack([ec[0][1], ec[1][1]])
        self.occ = np.zeros((2, self.nao)); self.occ[0, :nel[0]] = 1; self.occ[1, :nel[1]] = 1
        self._lpq_ao, self._h1e_ao = lpq, h1e
        self.e_nuc = self.mol.e_nuc
        self.e_elec = 0.5 * sum(np.sum(dm[s] * (h1e + f[s])) for s in range(2))
        self.e_tot = self.e_elec + self.e_nuc
        return self

    @staticmethod
    def _fock_ao(lpq, h1e, dm):
        j, _ = _jk(lpq, dm[0] + dm[1])
        return np.stack([h1e + j - _jk(lpq, dm[s])[1] for s in range(2)])

    @property
    def nelec(self):
        return self.nalph + self.nbeta

    @property
    def chempot(self):
        return tuple(0.5 * (self.e[s][n - 1] + self.e[s][n]) for s, n in enumerate((self.nalph, self.nbeta)))

    @property
    def h1e_mo(self):
        return np.stack([self.c[s].T @ self._h1e_ao @ self.c[s] for s in range(2)])

    @property
    def eri_mo(self):
        return np.stack([np.einsum("Qpq,pi,qj->Qij", self._lpq_ao, self.c[s], self.c[s], optimize=True)
                         for s in range(2)])

    @property
    def rdm1_mo(self):
        return np.stack([np.diag(self.occ[s]) for s in range(2)])

    def get_fock(self, rdm1, basis="mo"):
        cc = self.c if basis == "mo" else np.stack([np.eye(self.nao)] * 2)
        dm = np.stack([cc[s] @ rdm1[s] @ cc[s].T for s in range(2)])
        f = self._fock_ao(self._lpq_ao, self._h1e_ao, dm)
        return np.stack([cc[s].T @ f[s] @ cc[s] for s in range(2)])

    @staticmethod
    def energy_1body(h1e, rdm1, fock):
        return 0.5 * sum(np.sum(rdm1[s] * (h1e[s] + fock[s])) for s in range(2))


# ---- the same protocol from converged arrays (any SCF program) -----------------------------------------------
class ArraysRHF(ModelRHF):
    """Closed-shell ``hf`` object from a converged density-fitted SCF given as arrays: ``lpq_ao`` (naux, nao, nao) DF
    tensor, ``h1e_ao`` core Hamiltonian, ``mo_coeff`` (C^T S C = 1; the AO basis need not be orthonormal), ``mo_energy``,
    number of doubly occupied orbitals, nuclear repulsion and total energy."""

    def __init__(self, lpq_ao, h1e_ao, mo_coeff, mo_energy, nocc, e_nuc=0.0, e_tot=None, basis="arrays"):
        lpq_ao = np.asarray(lpq_ao, dtype=np.float64); h1e_ao = np.asarray(h1e_ao, dtype=np.float64)
        self.nao, self.nocc_hf, self.naux = h1e_ao.shape[0], int(nocc), lpq_ao.shape[0]
        assert lpq_ao.shape == (self.naux, self.nao, self.nao)
        self.mol = Molecule(e_nuc=e_nuc, basis=basis)
        self._lpq_ao, self._h1e_ao = lpq_ao, h1e_ao
        self.c = np.asarray(mo_coeff, dtype=np.float64)
        self.e = np.asarray(mo_energy, dtype=np.float64)
        self.occ = np.zeros(self.e.size); self.occ[:self.nocc_hf] = 2.0
        self.e_nuc = float(e_nuc)
        dm = 2 * self.c[:, :self.nocc_hf] @ self.c[:, :self.nocc_hf].T
        j, k = _jk(lpq_ao, dm)
        self.e_elec = 0.5 * np.sum(dm * (2 * h1e_ao + j - 0.5 * k))
        self.e_tot = float(e_tot) if e_tot is not None else self.e_elec + self.e_nuc
        self._done = True

    def run(self, *args, **kwargs):
        return self


class ArraysUHF(ModelUHF):
    """Open-shell counterpart of ``ArraysRHF``: ``mo_coeff``, ``mo_energy`` are (2, ...) stacks (alpha, beta)."""

    def __init__(self, lpq_ao, h1e_ao, mo_coeff, mo_energy, nalph, nbeta, e_nuc=0.0, e_tot=None, basis="arrays"):
        lpq_ao = np.asarray(lpq_ao, dtype=np.float64); h1e_ao = np.asarray(h1e_ao, dtype=np.float64)
        self.nao, self.nalph, self.nbeta, self.naux = h1e_ao.shape[0], int(nalph), int(nbeta), lpq_ao.shape[0]
        self.mol = Molecule(e_nuc=e_nuc, basis=basis)
        self._lpq_ao, self._h1e_ao = lpq_ao, h1e_ao
        self.c = np.asarray(mo_coeff, dtype=np.float64)
        self.e = np.asarray(mo_energy, dtype=np.float64)
        self.occ = np.zeros((2, self.e.shape[1])); self.occ[0, :self.nalph] = 1; self.occ[1, :self.nbeta] = 1
        self.e_nuc = float(e_nuc)
        dm = np.stack([self.c[s][:, :n] @ self.c[s][:, :n].T for s, n in enumerate((self.nalph, self.nbeta))])
        f = self._fock_ao(lpq_ao, h1e_ao, dm)
        self.e_elec = 0.5 * sum(np.sum(dm[s] * (h1e_ao + f[s])) for s in range(2))
        self.e_tot = float(e_tot) if e_tot is not None else self.e_elec + self.e_nuc

    def run(self, *args, **kwargs):
        return self


def from_pyscf(mf):
    """``hf`` object for the AGF2 drivers from a converged density-fitted ``pyscf.scf.RHF`` / ``UHF`` (the provider the
    reference wraps in auxgf/hf/hf.py:99-144): core Hamiltonian, MO coefficients / energies and the Cholesky-like DF
    tensor ``mf.with_df._cderi`` unpacked to (naux, nao, nao).  PySCF cannot be installed in the image this package is
    developed in, so this adapter is exercised only through ``ArraysRHF`` / ``ArraysUHF`` (tests/test_drivers_cpu.py)."""
    try:
        from pyscf import lib as pyscf_lib
    except ImportError as exc:         # pragma: no cover
        raise ImportError("from_pyscf needs PySCF; use ModelRHF / ModelUHF or ArraysRHF / ArraysUHF without it") from exc
    if getattr(mf, "with_df", None) is None:
        mf = mf.density_fit()
        mf.run()
    if mf.with_df._cderi is None:
        mf.with_df.build()
    lpq = pyscf_lib.unpack_tril(np.asarray(mf.with_df._cderi))
    h1e = mf.get_hcore()
    e_nuc = mf.mol.energy_nuc()
    mo_coeff, mo_energy = np.asarray(mf.mo_coeff), np.asarray(mf.mo_energy)
    if mo_coeff.ndim == 2:
        return ArraysRHF(lpq, h1e, mo_coeff, mo_energy, int(mf.mol.nelectron // 2), e_nuc, mf.e_tot, basis=mf.mol.basis)
    nalph, nbeta = mf.mol.nelec
    return ArraysUHF(lpq, h1e, mo_coeff, mo_energy, nalph, nbeta, e_nuc, mf.e_tot, basis=mf.mol.basis)
