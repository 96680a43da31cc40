"""CPU-only tests of the host-side driver logic (Aux container, OptRAGF2/OptUAGF2 control flow, Fock loop,
energies, multi-rank sharding over gloo).  The compute calls are routed to the oracle through
tests/oracle_backend.py -- this exercises the host code, NOT the CUDA path (that is tests/test_*_gpu.py)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle_backend import OracleBackend  # noqa: E402

from auxgf_b200 import backend  # noqa: E402
from auxgf_b200.aux import Aux, energy  # noqa: E402
from auxgf_b200.hf import ModelRHF, ModelUHF  # noqa: E402
from oracle import agf2_oracle as orc  # noqa: E402


@pytest.fixture()
def oracle_backend():
    prev = backend.use(OracleBackend())
    yield
    backend.use(prev)


def test_aux_container():
    rng = np.random.default_rng(0)
    e = np.sort(rng.standard_normal(7)); v = rng.standard_normal((3, 7))
    a = Aux(e, v, chempot=0.1)
    assert a.nphys == 3 and a.naux == 7 and a.nocc + a.nvir == 7
    occ, vir = a.as_occupied(), a.as_virtual()
    assert np.all(occ.e < 0.1) and np.all(vir.e >= 0.1) and occ.chempot == 0.1
    b = occ + vir
    assert b.naux == 7 and np.allclose(np.sort(b.e), e)
    h = rng.standard_normal((3, 3)); h = h + h.T
    m = a.as_hamiltonian(h, chempot=0.3)
    assert m.shape == (10, 10) and np.allclose(m, m.T) and np.allclose(np.diag(m)[3:], e - 0.3)
    buf = np.ones((10, 10))
    assert np.allclose(a.as_hamiltonian(h, chempot=0.3, out=buf), m)
    assert np.allclose(a.moment(1), (v * e) @ v.T) and a.moment([0, 1]).shape == (2, 3, 3)
    c = a.new(e[:2], v[:, :2])
    assert c.chempot == a.chempot and c.naux == 2
    with pytest.raises(ValueError):
        a + Aux(e, rng.standard_normal((4, 7)))
    with pytest.raises(ValueError):
        a.as_hamiltonian(np.eye(4))


def test_aux_dot_eq_and_io(tmp_path):
    rng = np.random.default_rng(3)
    e = np.sort(rng.standard_normal(6)); v = rng.standard_normal((4, 6))
    a = Aux(e, v, chempot=-0.2)
    h = rng.standard_normal((4, 4)); h = h + h.T
    x = rng.standard_normal((10, 3))
    assert np.allclose(a.dot(h, x, chempot=0.4), a.as_hamiltonian(h, chempot=0.4) @ x)
    assert np.allclose(a.dot(h, x[:, 0]), a.as_hamiltonian(h) @ x[:, 0])
    perm = rng.permutation(6)
    assert a == Aux(e[perm], v[:, perm], chempot=-0.2)          # same spectrum, poles in another order
    assert not (a == Aux(e + 0.05, v, chempot=-0.2))
    assert np.allclose(a.w_occ.sum() + a.w_vir.sum(), a.w.sum()) and a.memsize() > 0
    a.save(str(tmp_path / "se.pickle"))
    b = Aux.load(str(tmp_path / "se.pickle"))
    assert b == a and b.chempot == a.chempot
    # checkpoint format of the reference (auxgf/aux/aux.py:735-761: pickle of __dict__ = _ener, _coup, chempot):
    # what we write is what the reference's Aux.load setattr()s, and a reference-written file loads here
    import pickle
    with open(str(tmp_path / "se.pickle"), "rb") as f:
        d = pickle.load(f)
    assert set(d) == {"_ener", "_coup", "chempot"} and d["_coup"].shape == (4, 6)
    with open(str(tmp_path / "ref.pickle"), "wb") as f:
        pickle.dump({"_ener": e, "_coup": v, "chempot": -0.2}, f)
    assert Aux.load(str(tmp_path / "ref.pickle")) == a
    with open(str(tmp_path / "old.pickle"), "wb") as f:              # round-1 files
        pickle.dump({"e": e, "v": v, "chempot": -0.2}, f)
    assert Aux.load(str(tmp_path / "old.pickle")) == a


def test_checkpoint_option(oracle_backend, tmp_path, monkeypatch):
    """`checkpoint=True` (mpi/agf2/optragf2.py:35,568-570): rank 0 writes the density matrix and the self-energy."""
    from auxgf_b200.agf2 import OptRAGF2
    monkeypatch.chdir(tmp_path)
    hf = ModelRHF(nao=6, nocc=2, naux=10, seed=0).run()
    gf2 = OptRAGF2(hf, verbose=False, maxiter=2, checkpoint=True).run()
    assert np.allclose(np.loadtxt("rdm1_chk.dat"), gf2.rdm1)
    assert Aux.load("se_chk.pickle") == gf2.se


def test_unknown_option_rejected(oracle_backend):
    from auxgf_b200.agf2 import OptRAGF2
    hf = ModelRHF(nao=6, nocc=2, naux=10, seed=0).run()
    with pytest.raises(ValueError):
        OptRAGF2(hf, verbose=False, not_an_option=1)


def test_optragf2_host_logic(oracle_backend):
    from auxgf_b200.agf2 import OptRAGF2
    hf = ModelRHF(nao=10, nocc=3, naux=24, seed=1).run()
    gf2 = OptRAGF2(hf, verbose=False, etol=1e-9)
    # iteration 0: Fock matrix of the HF density is diag(e); se has 2*nphys poles; MP2 energy negative
    assert np.abs(gf2.get_fock() - np.diag(hf.e)).max() < 1e-10
    assert gf2.se.naux == 2 * hf.nao and gf2.nmom == (None, 0)
    assert -1.0 < gf2.e_mp2 < 0 and abs(gf2.e_tot - (hf.e_tot + gf2.e_mp2)) < 1e-12
    # the compressed self-energy reproduces the zeroth/first moments of the uncompressed pole list:
    # occupied sector of iteration 0 == explicit DF-MP2 poles (auxgf/aux/dfrmp2.py:82-109)
    l3 = orc.np_unpack_tril(gf2.eri, hf.nao)
    o, v = slice(0, 3), slice(3, 10)
    xija = np.einsum("Qxi,Qja->xija", l3[:, :, o], l3[:, o, v])
    eija = hf.e[o][:, None, None] + hf.e[o][None, :, None] - hf.e[v][None, None, :]
    w = 2 * xija - xija.transpose(0, 2, 1, 3)
    t0 = np.einsum("xija,yija->xy", xija, w)
    t1 = np.einsum("xija,ija,yija->xy", xija, eija, w)
    se_occ = gf2.se.as_occupied()
    assert np.abs(se_occ.moment(0) - t0).max() < 1e-10 and np.abs(se_occ.moment(1) - t1).max() < 1e-10
    gf2.run()
    assert gf2.converged and gf2.iteration < 30
    # IP/EA consistency (reference test: tests/agf2/test_optragf2.py:45-73)
    w_, v_ = gf2.se.eig(gf2.get_fock())
    assert abs(gf2.ip[0] + w_[w_ < gf2.chempot].max()) < 1e-8
    assert abs(gf2.ea[0] - w_[w_ >= gf2.chempot].min()) < 1e-8
    # electron number of the converged density
    assert abs(np.trace(gf2.rdm1) - hf.nelec) < 1e-5
    assert set(gf2._timings) >= {"setup", "build", "fock", "energy", "total"}


def test_optuagf2_host_logic(oracle_backend):
    from auxgf_b200.agf2 import OptRAGF2, OptUAGF2
    # closed-shell molecule treated unrestricted must reproduce the restricted MP2 energy
    rhf = ModelRHF(nao=8, nocc=3, naux=20, seed=4).run()
    uhf = ModelUHF(nao=8, nalph=3, nbeta=3, naux=20, seed=4).run()
    assert abs(rhf.e_tot - uhf.e_tot) < 1e-9
    r = OptRAGF2(rhf, verbose=False)
    u = OptUAGF2(uhf, verbose=False)
    assert abs(r.e_mp2 - u.e_mp2) < 1e-9
    # open shell
    uhf = ModelUHF(nao=8, nalph=3, nbeta=2, naux=20, seed=2).run()
    u = OptUAGF2(uhf, verbose=False, etol=1e-8)
    assert u.se[0].naux == 16 and u.se[1].naux == 16 and u.e_mp2 < 0
    u.run()
    assert u.converged
    assert abs(np.trace(u.rdm1[0]) - 3) < 1e-5 and abs(np.trace(u.rdm1[1]) - 2) < 1e-5


def test_scs_options_are_honoured(oracle_backend):
    from auxgf_b200.agf2 import OptRAGF2
    hf = ModelRHF(nao=8, nocc=3, naux=20, seed=5).run()
    a = OptRAGF2(hf, verbose=False)
    b = OptRAGF2(hf, verbose=False, os_factor=1.2, ss_factor=1.0 / 3.0)
    assert abs(a.e_mp2 - b.e_mp2) > 1e-6


def test_energy_functions_match_reference_loops(oracle_backend):
    """energy_2body_aux restated as a matrix product == the reference's per-pole loop (energy.py:43-49)."""
    rng = np.random.default_rng(3)
    n = 5
    gf = Aux(np.sort(rng.standard_normal(12)), rng.standard_normal((n, 12)), chempot=0.0)
    se_e = np.concatenate([-1 - rng.random(6), 1 + rng.random(7)])
    se = Aux(se_e, rng.standard_normal((n, 13)), chempot=0.0)
    ref = 0.0
    for l in range(gf.nocc):
        vxl = gf.v[:, l]; vxk = se.v[:, se.nocc:]
        ref += np.einsum("xk,yk,x,y,k->", vxk, vxk, vxl, vxl, 1.0 / (gf.e[l] - se.e[se.nocc:]))
    assert abs(energy.energy_2body_aux(gf, se) - 2 * ref) < 1e-12
    mo = np.sort(rng.standard_normal(n))
    occ = mo < 0
    ref = np.sum(se.v_vir[occ] ** 2 / (mo[occ][:, None] - se.e_vir[None]))
    assert abs(energy.energy_mp2_aux(mo, se) - ref) < 1e-12


class _GlooComm:
    """TEST-ONLY communicator (torch.distributed over gloo on CPU) with the interface of auxgf_b200.util.mpi."""

    def __init__(self, rank, size):
        self.rank, self.size = rank, size

    def reduce(self, m):
        import torch
        import torch.distributed as dist
        t = torch.from_numpy(np.ascontiguousarray(m, dtype=np.float64))
        dist.all_reduce(t)
        return t.numpy()

    def barrier(self):
        import torch.distributed as dist
        dist.barrier()


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle_backend import OracleBackend as OB
    from auxgf_b200 import backend as be
    from auxgf_b200.agf2 import OptRAGF2
    from auxgf_b200.util import mpi
    be.use(OB())
    mpi.use(_GlooComm(rank, world))
    hf = ModelRHF(nao=8, nocc=3, naux=20, seed=6).run()
    g = OptRAGF2(hf, verbose=False)
    g.options["maxiter"] = 2
    g.run()
    q.put((rank, g.e_mp2, g.e_tot))
    dist.destroy_process_group()


def test_two_rank_sharding_gloo(oracle_backend):
    """world_size = 2 over gloo: each rank builds its share of the moments, one all-reduce per sector;
    energies must equal the single-process run (replaces auxgf/mpi/agf2/optragf2.py:370-395)."""
    import torch.multiprocessing as mp
    from auxgf_b200.agf2 import OptRAGF2
    hf = ModelRHF(nao=8, nocc=3, naux=20, seed=6).run()
    g = OptRAGF2(hf, verbose=False)
    g.options["maxiter"] = 2
    g.run()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, e_mp2, e_tot in res:
        assert abs(e_mp2 - g.e_mp2) < 1e-10 and abs(e_tot - g.e_tot) < 1e-9


def test_hf_from_arrays_matches_model_and_is_basis_invariant(oracle_backend):
    """ArraysRHF / ArraysUHF (what from_pyscf builds): the same converged SCF handed over as arrays gives the same
    protocol quantities and the same AGF2 energies -- also in a non-orthonormal AO basis (C^T S C = 1 only), which is
    what a real SCF program provides."""
    from auxgf_b200.agf2 import OptRAGF2, OptUAGF2
    from auxgf_b200.hf import ArraysRHF, ArraysUHF
    m = ModelRHF(nao=9, nocc=3, naux=20, seed=5).run()
    rng = np.random.default_rng(0)
    x = np.eye(9) + 0.2 * rng.standard_normal((9, 9))            # AO basis change: phi' = phi X
    xi = np.linalg.inv(x)
    lpq = np.einsum("Qpq,pa,qb->Qab", m._lpq_ao, x, x)
    h1e = x.T @ m._h1e_ao @ x
    for a in (ArraysRHF(m._lpq_ao, m._h1e_ao, m.c, m.e, 3, m.e_nuc), ArraysRHF(lpq, h1e, xi @ m.c, m.e, 3, m.e_nuc)):
        assert abs(a.e_tot - m.e_tot) < 1e-10 and a.nelec == m.nelec and abs(a.chempot - m.chempot) < 1e-14
        assert np.abs(a.h1e_mo - m.h1e_mo).max() < 1e-10 and np.abs(a.eri_mo - m.eri_mo).max() < 1e-10
        assert np.abs(a.get_fock(m.rdm1_mo) - np.diag(m.e)).max() < 1e-9
        g, r = OptRAGF2(a, verbose=False, maxiter=2), OptRAGF2(m, verbose=False, maxiter=2)
        g.run(); r.run()
        assert abs(g.e_mp2 - r.e_mp2) < 1e-10 and abs(g.e_tot - r.e_tot) < 1e-8
    u = ModelUHF(nao=8, nalph=3, nbeta=2, naux=18, seed=6).run()
    au = ArraysUHF(u._lpq_ao, u._h1e_ao, u.c, u.e, 3, 2, u.e_nuc)
    assert abs(au.e_tot - u.e_tot) < 1e-10 and au.nelec == 5
    assert np.abs(au.eri_mo - u.eri_mo).max() < 1e-12 and np.abs(au.get_fock(u.rdm1_mo) - u.get_fock(u.rdm1_mo)).max() < 1e-12
    g, r = OptUAGF2(au, verbose=False, maxiter=1), OptUAGF2(u, verbose=False, maxiter=1)
    assert abs(g.e_mp2 - r.e_mp2) < 1e-10
