"""``OptUAGF2`` -- unrestricted DF-AGF2(None,0) on the B200; object API of ``auxgf/agf2/optuagf2.py``
(``build_part`` :147-233 returns ``(Aux_alpha, Aux_beta)``, ``get_fock`` :317-366, ``energy_mp2`` :270-279
averaged over spins).  Every quantity is an (alpha, beta) tuple / leading axis of length 2.  The beta
channel is the alpha code with the spin tuples reversed (optuagf2.py:230-231)."""
import numpy as np

from auxgf_b200 import backend
from auxgf_b200.aux import Aux, energy
from auxgf_b200.agf2 import optragf2
from auxgf_b200.agf2.fock import fock_loop_uhf
from auxgf_b200.agf2.optragf2 import OptRAGF2, pack_tril
from auxgf_b200.util import record_energy, record_time


class OptUAGF2(OptRAGF2):
    @record_time("setup")
    def _setup(self):
        hf = self.hf
        self.h1e = np.asarray(hf.h1e_mo, dtype=np.float64)
        eri = np.asarray(hf.eri_mo, dtype=np.float64)
        rdm1 = hf.rdm1_mo if self.options["dm0"] is None else self.options["dm0"]
        self.rdm1 = np.array(rdm1, dtype=np.float64)
        if self.rdm1.ndim == 2:
            self.rdm1 = np.stack([self.rdm1, self.rdm1])
        self.converged = False
        self.iteration = 0
        nact = hf.nao
        self.se = tuple(Aux(np.zeros(0), np.zeros((nact, 0)), chempot=hf.chempot[s]) for s in range(2))
        self.gf = tuple(self.se[s].new(hf.e[s], np.eye(nact)) for s in range(2))
        if eri.ndim == 4:
            eri = np.stack([pack_tril(eri[0]), pack_tril(eri[1])])
        self.eri = eri
        be = backend.get()
        self._eri_dev = tuple(be.stage_eri(eri[s], nact, sym_in="s2") for s in range(2))

    @staticmethod
    def build_part(gf_occ, gf_vir, eri, sym_in="s2", os_factor=1.0, ss_factor=1.0):
        be = backend.get()
        nphys = gf_occ[0].nphys
        staged = tuple(e if hasattr(e, "sector") else be.stage_eri(np.asarray(e), nphys, sym_in=sym_in) for e in eri)

        def _build_part(s):
            o = tuple(gf_occ[::s]); v = tuple(gf_vir[::s]); er = tuple(staged[::s])
            mom = er[0].moments_uhf(er[0], er[1], o, v, os_factor=os_factor, ss_factor=ss_factor)
            e, c = er[0].compress(mom)
            return o[0].new(e, c)

        return _build_part(1), _build_part(-1)

    @record_time("build")
    def build(self):
        self.solve_dyson()
        gf_occ = tuple(g.as_occupied() for g in self.gf)
        gf_vir = tuple(g.as_virtual() for g in self.gf)
        kw = dict(os_factor=self.options["os_factor"], ss_factor=self.options["ss_factor"])
        se_occ = self.build_part(gf_occ, gf_vir, self._eri_dev, **kw)
        se_vir = self.build_part(gf_vir, gf_occ, self._eri_dev, **kw)
        self.se = (se_occ[0] + se_vir[0], se_occ[1] + se_vir[1])

    def solve_dyson(self):
        fock = self.get_fock()
        gf = []
        for s in range(2):
            e, c = self.se[s].eig_phys(fock[s])
            gf.append(self.se[s].new(e, c))
        self.gf = tuple(gf)

    @record_time("fock")
    def fock_loop(self):
        opts = dict(self.options["_fock_loop"])
        opts["fock_func"] = self.get_fock
        opts["log"] = self._log if self.verbose else None
        se, rdm1, converged = fock_loop_uhf(self.se, self.hf, self.rdm1, **opts)
        self.solve_dyson()
        self._log("Fock loop %s.  Chemical potentials = %.6f %.6f"
                  % (("converged" if converged else "did not converge",) + tuple(self.chempot)))

    def get_fock(self, rdm1=None):
        """f_a = h_a + J_a + J_b - K_a, f_b = h_b + J_b + J_a - K_b (optuagf2.py:317-366), J/K on the device."""
        rdm1 = self.rdm1 if rdm1 is None else rdm1
        j_a, k_a = self._eri_dev[0].get_jk(rdm1[0])
        j_b, k_b = self._eri_dev[1].get_jk(rdm1[1])
        return np.stack([self.h1e[0] + j_a + j_b - k_a, self.h1e[1] + j_b + j_a - k_b])

    @record_time("energy")
    @record_energy("mp2")
    def energy_mp2(self):
        emp2 = 0.5 * (energy.energy_mp2_aux(self.hf.e[0], self.se[0]) + energy.energy_mp2_aux(self.hf.e[1], self.se[1]))
        self._log("E(mp2) = %.12f" % emp2)
        return emp2

    def _write_checkpoint(self):
        """mpi/agf2/optuagf2.py:479-483"""
        np.savetxt("rdm1_alph_chk.dat", self.rdm1[0])
        np.savetxt("rdm1_beta_chk.dat", self.rdm1[1])
        self.se[0].save("se_alph_chk.pickle")
        self.se[1].save("se_beta_chk.pickle")

    @property
    def chempot(self):
        return tuple(s.chempot for s in self.se)

    @property
    def nphys(self):
        return self.hf.nao

    @property
    def naux(self):
        return tuple(s.naux for s in self.se)

    @property
    def ip(self):
        best = (-np.inf, None)
        for gf in self.gf:
            occ = gf.as_occupied()
            arg = int(np.argmax(occ.e))
            if occ.e[arg] > best[0]:
                best = (occ.e[arg], occ.v[:, arg])
        return -best[0], best[1]

    @property
    def ea(self):
        best = (np.inf, None)
        for gf in self.gf:
            vir = gf.as_virtual()
            arg = int(np.argmin(vir.e))
            if vir.e[arg] < best[0]:
                best = (vir.e[arg], vir.v[:, arg])
        return best


optragf2.OptUAGF2 = OptUAGF2
