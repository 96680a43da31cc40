"""GPU: the OptRAGF2 / OptUAGF2 drivers on the CUDA backend (device-resident DF tensor, device ao2mo,
k1_form/k2_moment, cuSOLVER compression and extended-Fock eigensolves) against the same drivers on the
oracle backend.  Tolerances of BASELINE.json: moments 1e-10 relative, energies and IP/EA 1e-8 Ha."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from oracle_backend import OracleBackend  # noqa: E402

from auxgf_b200 import backend  # noqa: E402
from auxgf_b200.hf import ModelRHF, ModelUHF  # noqa: E402
from oracle import agf2_oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def _run(cls_name, hf, niter, **kw):
    from auxgf_b200 import agf2
    g = getattr(agf2, cls_name)(hf, verbose=False, **kw)
    e0 = g.e_mp2
    g.options["maxiter"] = niter
    g.run()
    return g, e0


def _with_oracle(fn):
    prev = backend.use(OracleBackend())
    try:
        return fn()
    finally:
        backend.use(prev)


def test_device_ao2mo_matches_oracle():
    from auxgf_b200.lib import agf2 as lib
    import ctypes
    import torch
    rng = np.random.default_rng(0)
    n, p, o, v = 37, 21, 6, 15                      # odd naux / nvir: padded pitches
    l3 = rng.standard_normal((n, p, p)); l3 = l3 + l3.transpose(0, 2, 1)
    il = np.tril_indices(p)
    packed = np.ascontiguousarray(l3[:, il[0], il[1]])
    co, cv = rng.standard_normal((p, o)), rng.standard_normal((p, v))
    eri = lib.DeviceERI(packed, p)
    ixq_p, ldi, qja_p, ldq = eri.sector(co, cv)
    torch.cuda.synchronize()
    assert ldi >= n and ldq >= v and ldi % 2 == 0 and ldq % 2 == 0

    def fetch(ptr, count):
        buf = np.empty(count)
        rc = ctypes.CDLL("libcudart.so.12").cudaMemcpy(ctypes.c_void_p(buf.ctypes.data), ctypes.c_void_p(ptr),
                                                       ctypes.c_size_t(count * 8), 2)
        assert rc == 0
        return buf
    ixq = fetch(ixq_p, o * p * ldi).reshape(o, p, ldi)[:, :, :n]
    qja = fetch(qja_p, n * o * ldq).reshape(n, o, ldq)[:, :, :v]
    ref_ixq = orc.np_ao2mo(packed, co, np.eye(p)).reshape(n, o, p).transpose(1, 2, 0)
    ref_qja = orc.np_ao2mo(packed, co, cv).reshape(n, o, v)
    assert relerr(ixq, ref_ixq) < 1e-12 and relerr(qja, ref_qja) < 1e-12
    eri.close()


def test_device_jk_matches_oracle():
    from auxgf_b200.lib import agf2 as lib
    rng = np.random.default_rng(1)
    n, p = 45, 33
    l3 = rng.standard_normal((n, p, p)); l3 = l3 + l3.transpose(0, 2, 1)
    d = rng.standard_normal((p, p)); d = d + d.T
    eri = lib.DeviceERI(l3, p, packed=False)
    j, k = eri.get_jk(d)
    rj = np.einsum("Qpq,Qrs,rs->pq", l3, l3, d, optimize=True)
    rk = np.einsum("Qpr,rs,Qqs->pq", l3, d, l3, optimize=True)
    assert relerr(j, rj) < 1e-12 and relerr(k, rk) < 1e-12
    eri.close()


def test_optragf2_cuda_vs_oracle():
    hf = ModelRHF(nao=24, nocc=5, naux=70, seed=1).run()
    g, e0 = _run("OptRAGF2", hf, 3)
    r, r0 = _with_oracle(lambda: _run("OptRAGF2", hf, 3))
    assert abs(e0 - r0) < 1e-8 and abs(g.e_tot - r.e_tot) < 1e-8
    assert abs(g.e_1body - r.e_1body) < 1e-8 and abs(g.e_2body - r.e_2body) < 1e-8
    assert abs(g.ip[0] - r.ip[0]) < 1e-8 and abs(g.ea[0] - r.ea[0]) < 1e-8
    assert abs(g.chempot - r.chempot) < 1e-7
    # moments: 1e-10 wherever both drivers start from the same state -- the construction-time (MP2) build, and one
    # build() from the oracle driver's final state; after independent Fock loops (dtol = 1e-8) only to 1e-8
    g0 = _run("OptRAGF2", hf, 0)[0]
    r0 = _with_oracle(lambda: _run("OptRAGF2", hf, 0)[0])
    for k in (0, 1):
        assert relerr(g0.se.moment(k), r0.se.moment(k)) < 1e-10
        assert relerr(g.se.moment(k), r.se.moment(k)) < 1e-8
    g.se, g.gf, g.rdm1 = r.se.copy(), r.gf.copy(), r.rdm1.copy()
    g.build()
    _with_oracle(r.build)
    for k in (0, 1):
        assert relerr(g.se.moment(k), r.se.moment(k)) < 1e-10


def test_build_part_moments_cuda_vs_oracle():
    """One sector through the public static build_part: moments of the returned Aux vs the oracle (1e-10)."""
    from auxgf_b200.agf2 import OptRAGF2
    hf = ModelRHF(nao=40, nocc=9, naux=130, seed=3).run()
    g = OptRAGF2(hf, verbose=False)
    gf_occ, gf_vir = g.gf.as_occupied(), g.gf.as_virtual()
    se = OptRAGF2.build_part(gf_occ, gf_vir, g.eri)          # host eri -> staged on the fly
    ref = _with_oracle(lambda: OptRAGF2.build_part(gf_occ, gf_vir, g.eri))
    for k in (0, 1):
        assert relerr(se.moment(k), ref.moment(k)) < 1e-10
    np.testing.assert_allclose(np.sort(se.e), np.sort(ref.e), rtol=0, atol=1e-9)
    se2 = OptRAGF2.build_part(gf_vir, gf_occ, g._eri_dev, os_factor=1.2, ss_factor=1.0 / 3.0)
    ref2 = _with_oracle(lambda: OptRAGF2.build_part(gf_vir, gf_occ, g.eri, os_factor=1.2, ss_factor=1.0 / 3.0))
    for k in (0, 1):
        assert relerr(se2.moment(k), ref2.moment(k)) < 1e-10


def test_optuagf2_cuda_vs_oracle():
    uhf = ModelUHF(nao=16, nalph=4, nbeta=3, naux=50, seed=2).run()
    g, e0 = _run("OptUAGF2", uhf, 2)
    r, r0 = _with_oracle(lambda: _run("OptUAGF2", uhf, 2))
    assert abs(e0 - r0) < 1e-8 and abs(g.e_tot - r.e_tot) < 1e-8
    assert abs(g.ip[0] - r.ip[0]) < 1e-8 and abs(g.ea[0] - r.ea[0]) < 1e-8
    g0 = _run("OptUAGF2", uhf, 0)[0]
    r0 = _with_oracle(lambda: _run("OptUAGF2", uhf, 0)[0])
    for s in range(2):
        for k in (0, 1):
            assert relerr(g0.se[s].moment(k), r0.se[s].moment(k)) < 1e-10      # same inputs: north_star tolerance
            assert relerr(g.se[s].moment(k), r.se[s].moment(k)) < 1e-8         # after independent Fock loops
    g.se = tuple(x.copy() for x in r.se); g.gf = tuple(x.copy() for x in r.gf); g.rdm1 = r.rdm1.copy()
    g.build()
    _with_oracle(r.build)
    for s in range(2):
        for k in (0, 1):
            assert relerr(g.se[s].moment(k), r.se[s].moment(k)) < 1e-10


def test_fock_solver_and_energy_on_device():
    """lib.FockSolver (resident couplings, device-assembled extended Fock matrix, cuSOLVER dsyevd, physical rows only)
    against np.linalg.eigh; the chemical-potential gradient from physical rows only against the reference formula
    with the auxiliary rows (chempot.py:77-118); energy_2body on the device against the reference loop."""
    from auxgf_b200.lib import agf2 as lib
    from auxgf_b200.aux import Aux, energy
    from auxgf_b200.agf2 import chempot as cp
    rng = np.random.default_rng(7)
    n, m = 23, 61
    f = rng.standard_normal((n, n)); f = f + f.T
    v = rng.standard_normal((n, m)) * 0.3
    e = np.concatenate([-1 - rng.random(30), 1 + rng.random(31)])
    h = np.zeros((n + m, n + m)); h[:n, :n] = f; h[:n, n:] = v; h[n:, :n] = v.T
    s = lib.FockSolver()
    s.set_couplings(v)
    for shift in (0.0, 0.37):
        h[n:, n:] = np.diag(e - shift)
        w, vp = s.solve(f, e, shift)
        wr, vr = np.linalg.eigh(h)
        np.testing.assert_allclose(w, wr, rtol=0, atol=1e-11)
        # eigenvectors up to sign: compare the projectors onto the physical space, v_phys diag(w) v_phys^T etc.
        assert relerr((vp * w) @ vp.T, (vr[:n] * wr) @ vr[:n].T) < 1e-11
        assert relerr(vp @ vp.T, np.eye(n)) < 1e-11
    s2 = lib.FockSolver(); s2.set_couplings(v[:, :40])        # two handles in flight (alpha / beta)
    s.solve_async(f, e, 0.1); s2.solve_async(f, e[:40], 0.2)
    w1, _ = s.wait(); w2, _ = s2.wait()
    h[n:, n:] = np.diag(e - 0.1)
    np.testing.assert_allclose(w1, np.linalg.eigvalsh(h), rtol=0, atol=1e-11)
    h2 = h[:n + 40, :n + 40].copy(); h2[n:, n:] = np.diag(e[:40] - 0.2)
    np.testing.assert_allclose(w2, np.linalg.eigvalsh(h2), rtol=0, atol=1e-11)
    s.close(); s2.close()
    # gradient of the electron-number error: physical rows only == the reference's auxiliary-row formula
    se = Aux(e, v, chempot=0.0)
    x = 0.05
    g = cp.gradient(x, se, f, 20, return_val=True)
    h[n:, n:] = np.diag(e - x)
    wr, vr = np.linalg.eigh(h)
    from auxgf_b200.util import find_chempot
    mu, err = find_chempot(n, 20, (wr, vr))
    nocc = int(np.sum(wr < mu))
    h1 = -(vr[n:, nocc:].T @ vr[n:, :nocc])
    z = -h1 / (wr[None, :nocc] - wr[nocc:, None])
    rdm1 = (vr[:n, :nocc] @ (vr[:n, nocc:] @ z).T) * 4
    assert abs(g[0] - err ** 2) < 1e-10 and abs(g[1] - 2 * err * np.trace(rdm1)) < 1e-9 * max(1.0, abs(g[1]))
    # two-body energy: device GEMM + reduction vs the reference's per-pole loop (energy.py:43-49)
    gf = Aux(np.sort(rng.standard_normal(70)), rng.standard_normal((n, 70)), chempot=0.0)
    ref = 0.0
    for l in range(gf.nocc):
        vxl = gf.v[:, l]; vxk = se.v[:, se.nocc:]
        ref += np.einsum("xk,yk,x,y,k->", vxk, vxk, vxl, vxl, 1.0 / (gf.e[l] - se.e[se.nocc:]))
    assert abs(energy.energy_2body_aux(gf, se) - 2 * ref) < 1e-10 * max(1.0, abs(ref))
