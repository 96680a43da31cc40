"""Compute backend of the drivers.

The package's only backend is the CUDA one (``CudaBackend``: every call goes through the ctypes C-ABI of
``libagf2_b200.so``).  There is no CPU fallback and nothing here ever selects one: if the library or the
GPU is missing the calls raise.  ``use()`` exists so that the CPU-only test-suite can inject its oracle
(``tests/oracle_backend.py``) to exercise the host-side driver logic; the package itself never calls it.
"""
import numpy as np


class CudaBackend:
    """Moment build, pole compression and dense eigensolves on the B200."""
    name = "cuda"

    def __init__(self):
        self._lib = None

    @property
    def lib(self):
        if self._lib is None:
            from auxgf_b200.lib import agf2 as lib   # raises RuntimeError if the .so is missing
            self._lib = lib
        return self._lib

    # -- hot path ---------------------------------------------------------------------------
    def build_part_loop_rhf(self, ixq, qja, gf_occ, gf_vir, istart, iend, vv=None, vev=None,
                            os_factor=1.0, ss_factor=1.0):
        return self.lib.build_part_loop_rhf(ixq, qja, gf_occ, gf_vir, istart, iend, vv=vv, vev=vev,
                                            os_factor=os_factor, ss_factor=ss_factor)

    def build_part_loop_uhf(self, ixq, qja, gf_occ, gf_vir, istart, iend, vv=None, vev=None,
                            os_factor=1.0, ss_factor=1.0):
        return self.lib.build_part_loop_uhf(ixq, qja, gf_occ, gf_vir, istart, iend, vv=vv, vev=vev,
                                            os_factor=os_factor, ss_factor=ss_factor)

    def compress(self, vv, vev):
        return self.lib.compress(vv, vev)

    def eigh(self, a):
        return self.lib.eigh(np.ascontiguousarray(a, dtype=np.float64))

    def fock_solver(self):
        """Extended-Fock eigensolver with the couplings resident in HBM (``lib.FockSolver``)."""
        return self.lib.FockSolver()

    def energy_2body(self, gv, ge, sv, se):
        return self.lib.energy_2body(gv, ge, sv, se)

    # device-resident DF tensor + fused ao2mo/moment build (see auxgf_b200.lib.agf2.DeviceERI)
    def stage_eri(self, eri, nphys, sym_in="s2"):
        return self.lib.DeviceERI(eri, nphys, packed=(sym_in == "s2"))


_active = CudaBackend()


def get():
    return _active


def use(backend):
    """Install another backend object (test-suite only).  Returns the previous one."""
    global _active
    prev, _active = _active, backend
    return prev
