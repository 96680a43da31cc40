"""``Aux`` -- the pole container exchanged across the hot path (mirrors the interface of
``auxgf/aux/aux.py:16-921`` that ``OptRAGF2``/``OptUAGF2`` use: attributes ``e``, ``v``, ``chempot``,
``nphys``, ``naux``, ``nocc``, ``nvir``, ``e_occ``..., methods ``new``, ``copy``, ``as_occupied``,
``as_virtual``, ``as_hamiltonian``, ``eig``, ``moment``, ``__add__``, ``save``/``load``).

Energies ``e`` are an (naux,) float64 C array, couplings ``v`` an (nphys, naux) float64 C array.
The dense eigensolve of the extended Hamiltonian (``auxgf/aux/eig.py:23-33``) runs on the GPU through
``auxgf_b200.backend`` (cuSOLVER dsyevd)."""
import pickle

import numpy as np

from auxgf_b200 import backend


class Aux:
    def __init__(self, e, v, chempot=0.0):
        e = np.asarray([] if e is None else e, dtype=np.float64).reshape(-1)
        v = np.asarray([[]] if v is None else v, dtype=np.float64)
        if v.ndim != 2:
            v = v.reshape(v.shape[0] if v.ndim else 0, -1)
        if v.shape[1] != e.size:
            raise ValueError("couplings (%s) and energies (%d) do not match" % (v.shape, e.size))
        self._ener = np.ascontiguousarray(e)
        self._coup = np.ascontiguousarray(v)
        self.chempot = float(chempot)

    # ---- construction ------------------------------------------------------------------
    def new(self, e, v, chempot=None):
        """New poles inheriting the chemical potential (aux.py:781)."""
        return Aux(e, v, chempot=self.chempot if chempot is None else chempot)

    def copy(self):
        return Aux(self._ener.copy(), self._coup.copy(), chempot=self.chempot)

    def __add__(self, other):
        """Concatenate poles; chempot from the left operand (aux.py:811)."""
        if self.nphys != other.nphys:
            raise ValueError("Cannot combine auxiliaries with different physical space dimensions.")
        return Aux(np.concatenate((self.e, other.e)), np.concatenate((self.v, other.v), axis=1),
                   chempot=self.chempot)

    def save(self, filename):
        """Checkpoint in the reference's format: a pickle of the instance dict, keys ``_ener``, ``_coup``,
        ``chempot`` (aux.py:735-741), so that either code restarts from the other's ``se_chk.pickle``."""
        with open(filename, "wb") as f:
            pickle.dump({"_ener": self._ener, "_coup": self._coup, "chempot": self.chempot}, f)

    @classmethod
    def load(cls, filename):
        """Reads the reference's checkpoints (aux.py:744-761) and the ``e``/``v`` keyed ones of earlier versions."""
        with open(filename, "rb") as f:
            d = pickle.load(f)
        e = d["_ener"] if "_ener" in d else d["e"]
        v = d["_coup"] if "_coup" in d else d["v"]
        return cls(e, v, chempot=d.get("chempot", 0.0))

    # ---- views ---------------------------------------------------------------------------
    def as_occupied(self):
        m = self.e < self.chempot
        return Aux(self.e[m], self.v[:, m], chempot=self.chempot)

    def as_virtual(self):
        m = self.e >= self.chempot
        return Aux(self.e[m], self.v[:, m], chempot=self.chempot)

    def as_window(self, e_min, e_max):
        m = (self.e >= e_min) & (self.e < e_max)
        return Aux(self.e[m], self.v[:, m], chempot=self.chempot)

    def sort(self):
        idx = np.argsort(self.e, kind="stable")
        self._ener = np.ascontiguousarray(self._ener[idx])
        self._coup = np.ascontiguousarray(self._coup[:, idx])

    def as_hamiltonian(self, h_phys, chempot=0.0, out=None):
        """[[h_phys, v], [v^T, diag(e - chempot)]]  (aux.py:319-359)."""
        h_phys = np.asarray(h_phys)
        n, m = self.nphys, self.naux
        if h_phys.shape != (n, n):
            raise ValueError("physical space of h_phys and couplings must match.")
        if out is None:
            out = np.zeros((n + m, n + m), dtype=np.float64)
        else:
            out[n:, n:] = 0.0
        out[:n, :n] = h_phys
        out[:n, n:] = self.v
        out[n:, :n] = self.v.T
        out[n:, n:][np.diag_indices(m)] = self.e - chempot
        return out

    def eig(self, h_phys, chempot=0.0, out=None):
        """Eigenpairs of the extended Hamiltonian (aux.py:468 -> eig.py:23-33), on the device."""
        h_ext = self.as_hamiltonian(h_phys, chempot=chempot, out=out)
        return backend.get().eigh(h_ext)

    def eig_phys(self, h_phys, chempot=0.0, out=None):
        """(w, v[:nphys]) of the extended Hamiltonian: eigenvalues and the physical rows of the eigenvectors -- all
        that the Dyson equation, the density matrix and the chemical-potential gradient need (method.py:134-147,
        fock.py:119-143, chempot.py:77-118).  The couplings are uploaded once per ``Aux`` (they stay in HBM while the
        Fock loop re-diagonalises with new ``h_phys`` / ``chempot``); the matrix is assembled on the device and only
        w and the nphys rows come back."""
        solver = self._fock_solver()
        return solver.solve(np.asarray(h_phys, dtype=np.float64), self._ener, shift=chempot)

    def _fock_solver(self):
        solver = getattr(self, "_solver", None)
        if solver is None:
            solver = self._solver = backend.get().fock_solver()
            self._solver_for = None
        if self._solver_for is not self._coup:      # couplings replaced (sort, new poles): upload again
            solver.set_couplings(self._coup)
            self._solver_for = self._coup
        return solver

    def dot(self, h_phys, vec, chempot=0.0):
        """``as_hamiltonian(h_phys, chempot) @ vec`` without building the extended matrix (aux.py:420-465)."""
        h_phys = np.asarray(h_phys)
        n = self.nphys
        if h_phys.shape != (n, n):
            raise ValueError("physical space of h_phys and couplings must match.")
        vec = np.asarray(vec)
        shape = vec.shape
        vec = vec.reshape(n + self.naux, -1)
        out = np.empty(vec.shape, dtype=np.result_type(self.v.dtype, vec.dtype))
        out[:n] = h_phys @ vec[:n] + self.v @ vec[n:]
        out[n:] = self.v.T @ vec[:n] + (self.e[:, None] - chempot) * vec[n:]
        return out.reshape(shape)

    def memsize(self):
        """Approximate size in GB (aux.py:764-775)."""
        return (self.e.nbytes + self.v.nbytes + 8) / 1e9

    def __eq__(self, other):
        """Same spectrum: the reference compares spectral functions on an imaginary-frequency grid (aux.py:849-868,
        ImFqGrid(32, beta=8), Feynman ordering); restated here as  v (i w_n - e)^-1 v^T  on those Matsubara points."""
        if not isinstance(other, Aux):
            return NotImplemented
        if self.nphys != other.nphys:
            return False
        wn = (2 * np.arange(32) + 1) * np.pi / 8.0

        def spectrum(a):
            d = 1.0 / (1j * wn[:, None] - (a.e - a.chempot)[None, :])
            return np.einsum("xk,wk,yk->wxy", a.v, d, a.v)
        return bool(np.allclose(spectrum(self), spectrum(other)))

    __hash__ = None

    def moment(self, n):
        n = np.asarray(n, dtype=np.int64).reshape(-1)
        moms = np.stack([(self.v * self.e[None] ** k) @ self.v.T for k in n])
        return np.squeeze(moms)

    # ---- properties ------------------------------------------------------------------------
    @property
    def e(self):
        return self._ener

    @property
    def v(self):
        return self._coup

    @property
    def e_occ(self):
        return self.e[self.e < self.chempot]

    @property
    def e_vir(self):
        return self.e[self.e >= self.chempot]

    @property
    def v_occ(self):
        return self.v[:, self.e < self.chempot]

    @property
    def v_vir(self):
        return self.v[:, self.e >= self.chempot]

    @property
    def w(self):
        return np.einsum("xk,xk->k", self.v, self.v)

    @property
    def w_occ(self):
        return np.einsum("xk,xk->k", self.v_occ, self.v_occ)

    @property
    def w_vir(self):
        return np.einsum("xk,xk->k", self.v_vir, self.v_vir)

    @property
    def nphys(self):
        return self.v.shape[0]

    @property
    def naux(self):
        return self.v.shape[1]

    @property
    def nocc(self):
        return int(np.count_nonzero(self.e < self.chempot))

    @property
    def nvir(self):
        return int(np.count_nonzero(self.e >= self.chempot))

    def __repr__(self):
        return "Aux(nphys=%d, naux=%d, chempot=%.6f)" % (self.nphys, self.naux, self.chempot)
