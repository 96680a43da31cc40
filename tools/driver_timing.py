"""Whole-driver timing: OptRAGF2 on a seeded model molecule of benzene cc-pVTZ size (BASELINE.json configs[1]:
nphys=264, nocc=21, naux=700) for a few AGF2 iterations on the GPU; prints the per-phase wall times the reference
records with @record_time (auxgf/util/decor.py:5-13) and the shapes of the sector calls."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from auxgf_b200.agf2 import OptRAGF2
from auxgf_b200.hf import ModelRHF
from auxgf_b200.lib import agf2 as lib

nao, nocc, naux = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "264,21,700").split(",")]
niter = int(sys.argv[2]) if len(sys.argv) > 2 else 3
t0 = time.perf_counter()
# (weak coupling so that the plain fixed-point SCF of the model converges at this size)
hf = ModelRHF(nao=nao, nocc=nocc, naux=naux, seed=7, strength=0.1).run(tol=1e-9, maxiter=150)
print("model RHF (host NumPy): %.1f s, E = %.10f" % (time.perf_counter() - t0, hf.e_tot))
t0 = time.perf_counter()
gf2 = OptRAGF2(hf, verbose=False, maxiter=niter)
print("setup + MP2 build: %.2f s, E(mp2) = %.10f" % (time.perf_counter() - t0, gf2.e_mp2))
# count / time the two device services the Fock loop leans on (dense eigensolve, DF-JK)
stats = {"eigh": [0, 0.0, 0], "get_jk": [0, 0.0, 0], "fock_solve": [0, 0.0, 0]}
_eigh, _jk = lib.eigh, lib.DeviceERI.get_jk
_fs_async, _fs_wait = lib.FockSolver.solve_async, lib.FockSolver.wait


def fs_async_timed(self, fock, e, shift=0.0):
    self._t0 = time.perf_counter()
    return _fs_async(self, fock, e, shift)


def fs_wait_timed(self):
    r = _fs_wait(self)
    stats["fock_solve"][0] += 1; stats["fock_solve"][1] += time.perf_counter() - self._t0
    stats["fock_solve"][2] = max(stats["fock_solve"][2], self.nphys + self.naux)
    return r


def eigh_timed(a):
    t = time.perf_counter(); r = _eigh(a); stats["eigh"][0] += 1; stats["eigh"][1] += time.perf_counter() - t
    stats["eigh"][2] = max(stats["eigh"][2], a.shape[0])
    return r


def jk_timed(self, rdm1):
    t = time.perf_counter(); r = _jk(self, rdm1); stats["get_jk"][0] += 1; stats["get_jk"][1] += time.perf_counter() - t
    return r


lib.eigh = eigh_timed
lib.DeviceERI.get_jk = jk_timed
lib.FockSolver.solve_async = fs_async_timed
lib.FockSolver.wait = fs_wait_timed
lib.launch_count(reset=True)
t0 = time.perf_counter()
gf2.run()
dt = time.perf_counter() - t0
print("%d AGF2 iterations: %.2f s, E(tot) = %.10f, IP = %.6f, EA = %.6f, gf poles occ/vir = %d/%d"
      % (gf2.iteration, dt, gf2.e_tot, gf2.ip[0], gf2.ea[0], gf2.gf.as_occupied().naux, gf2.gf.as_virtual().naux))
print("phase timings (s):", {k: round(v, 3) for k, v in gf2._timings.items()})
print("kernel launches:", lib.launch_count())
print("eigh (host in/out): %d calls, %.2f s total (largest n = %d); extended-Fock solves with resident couplings "
      "(agf2_b200_fock_solve): %d calls, %.2f s total = %.2f ms each (largest n = %d); get_jk: %d calls, %.2f s total"
      % (stats["eigh"][0], stats["eigh"][1], stats["eigh"][2], stats["fock_solve"][0], stats["fock_solve"][1],
         1e3 * stats["fock_solve"][1] / max(stats["fock_solve"][0], 1), stats["fock_solve"][2], stats["get_jk"][0],
         stats["get_jk"][1]))
