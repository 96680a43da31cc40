"""TEST-ONLY backend: routes the drivers' compute calls to the CPU oracle (oracle/agf2_oracle.py) so that
the host-side driver logic (Aux, Fock loop, energies, multi-rank sharding) can be exercised without a
GPU, and so that GPU runs can be compared driver-to-driver.  Never imported by the package."""
import numpy as np

from auxgf_b200.util import mpi
from oracle import agf2_oracle as orc


class _HostERI:
    def __init__(self, eri, nphys, packed):
        eri = np.asarray(eri, dtype=np.float64)
        self.nphys = nphys
        self.full = orc.np_unpack_tril(eri, nphys) if packed else eri.reshape(-1, nphys, nphys)
        self.naux = self.full.shape[0]

    def _blocks(self, v_occ, v_vir):
        ixq = np.einsum("Qxp,pi->ixQ", self.full, v_occ, optimize=True)
        qja = np.einsum("jxQ,xa->Qja", ixq, v_vir, optimize=True)
        return ixq, qja

    def sector(self, v_occ, v_vir):
        return self._blocks(v_occ, v_vir)

    def get_jk(self, rdm1):
        """optragf2.py:358-398 restated with einsum."""
        rho = np.einsum("Qpq,pq->Q", self.full, rdm1)
        j = np.einsum("Q,Qpq->pq", rho, self.full)
        k = np.einsum("Qpr,rs,Qqs->pq", self.full, rdm1, self.full, optimize=True)
        return j, k

    def moments_rhf(self, gf_occ, gf_vir, os_factor=1.0, ss_factor=1.0):
        o, v, p = gf_occ.naux, gf_vir.naux, self.nphys
        part, nparts = mpi.rank(), mpi.size()
        ixq, qja = self._blocks(gf_occ.v, gf_vir.v)
        # the oracle shards by contiguous pole slices (auxgf/mpi/agf2/optragf2.py:370-374)
        lo, hi = (o * part) // nparts, (o * (part + 1)) // nparts
        vv, vev = orc.np_build_part_loop_rhf(ixq.reshape(o * p, -1), qja.reshape(self.naux, -1), gf_occ.e, gf_vir.e,
                                             p, o, v, self.naux, lo, hi, os_factor, ss_factor)
        return _Packed(mpi.reduce(np.stack([vv, vev])))

    @staticmethod
    def moments_uhf(eri_a, eri_b, gf_occ, gf_vir, os_factor=1.0, ss_factor=1.0):
        p = eri_a.nphys
        part, nparts = mpi.rank(), mpi.size()
        oa, ob, va, vb = gf_occ[0].naux, gf_occ[1].naux, gf_vir[0].naux, gf_vir[1].naux
        ixq_a, qja_a = eri_a._blocks(gf_occ[0].v, gf_vir[0].v)
        _, qja_b = eri_b._blocks(gf_occ[1].v, gf_vir[1].v)
        lo, hi = (oa * part) // nparts, (oa * (part + 1)) // nparts
        vv, vev = orc.np_build_part_loop_uhf(ixq_a.reshape(oa * p, -1), qja_a.reshape(eri_a.naux, -1),
                                             qja_b.reshape(eri_a.naux, -1), gf_occ[0].e, gf_occ[1].e, gf_vir[0].e,
                                             gf_vir[1].e, p, oa, ob, va, vb, eri_a.naux, lo, hi, os_factor, ss_factor)
        return _Packed(mpi.reduce(np.stack([vv, vev])))

    @staticmethod
    def compress(mom):
        return orc.np_compress(mom.a[0], mom.a[1])


class _Packed:
    """Host stand-in for the packed [vv|vev] CUDA tensor (summed over the ranks through auxgf_b200.util.mpi)."""
    def __init__(self, a):
        self.a = a


class _HostFockSolver:
    """Host stand-in for lib.FockSolver: dense np.linalg.eigh of the extended matrix (auxgf/aux/eig.py:23-33)."""

    def set_couplings(self, v):
        self._v = np.asarray(v)

    def solve(self, fock, e, shift=0.0):
        n, m = self._v.shape
        h = np.zeros((n + m, n + m))
        h[:n, :n] = fock
        h[:n, n:] = self._v
        h[n:, :n] = self._v.T
        h[n:, n:][np.diag_indices(m)] = np.asarray(e) - shift
        w, v = np.linalg.eigh(h)
        return w, v[:n]


class OracleBackend:
    name = "oracle"

    def fock_solver(self):
        return _HostFockSolver()

    def energy_2body(self, gv, ge, sv, se):
        """The reference's loop over occupied poles (auxgf/aux/energy.py:43-49)."""
        out = 0.0
        for l in range(gv.shape[1]):
            out += np.einsum("xk,yk,x,y,k->", sv, sv, gv[:, l], gv[:, l], 1.0 / (ge[l] - se))
        return float(out)

    def build_part_loop_rhf(self, ixq, qja, gf_occ, gf_vir, istart, iend, vv=None, vev=None, os_factor=1.0, ss_factor=1.0):
        o, v, p = gf_occ.naux, gf_vir.naux, gf_occ.nphys
        ixq = np.asarray(ixq).reshape(o * p, -1)
        return orc.np_build_part_loop_rhf(ixq, np.asarray(qja).reshape(ixq.shape[1], -1), gf_occ.e, gf_vir.e, p, o, v,
                                          ixq.shape[1], istart, iend, os_factor, ss_factor, vv, vev)

    def compress(self, vv, vev):
        return orc.np_compress(vv, vev)

    def eigh(self, a):
        return np.linalg.eigh(a)

    def stage_eri(self, eri, nphys, sym_in="s2"):
        return _HostERI(eri, nphys, packed=(sym_in == "s2"))
