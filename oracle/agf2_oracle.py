"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the DF-AGF2(None,0) moment build.

Nothing in the product path (``auxgf_b200/``) imports this module.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may use
it, and only as the checker / the reported CPU baseline.

Three independent statements of the same arithmetic live here (citations relative to
``/root/reference``):

* ``np_build_part_loop_rhf/uhf`` -- NumPy restatement of ``auxgf/lib/libagf2/optragf2.c:117-295``
  and ``optuagf2.c:125-315`` (equivalently the dead-code NumPy fallback in
  ``auxgf/agf2/optragf2.py:252-268`` / ``optuagf2.py:189-216``), generalised with the ``os``/``ss``
  factors of ``auxgf/aux/dfrmp2.py:82-84`` / ``dfump2.py:99-100``.
* ``c_build_part_loop_rhf/uhf`` -- ctypes binding of ``oracle/agf2_oracle.c`` (naive C loops).
* ``ref_build_part_loop_rhf/uhf`` -- ctypes binding of ``oracle/_ref/agf2.so``: the reference's own
  UNMODIFIED C loop compiled where it lies under ``/root/reference`` by ``oracle/Makefile``
  (symbols ``build_part_loop_rhf`` optragf2.c:32-43, ``build_part_loop_uhf`` optuagf2.c:31-47).

Parity pin: ``tests/test_oracle.py`` checks the three against each other and against the golden
vectors in ``tests/golden/`` (generated from ``oracle/_ref/agf2.so`` by ``oracle/gen_golden.py``).

Also restated here for the "next" rows: ``np_compress`` (optragf2.py:270-278), ``np_ao2mo``
(optragf2.py:175-214 semantics of ``pyscf.ao2mo._ao2mo.nr_e2`` 's2'->'s1').
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C,A")
_f64pw = np.ctypeslib.ndpointer(dtype=np.float64, flags="C,A,W")
_u32 = ctypes.c_uint32
_i64 = ctypes.c_int64
_dbl = ctypes.c_double


# --------------------------------------------------------------------------------------------
# NumPy restatement
# --------------------------------------------------------------------------------------------
def np_build_part_loop_rhf(ixq, qja, ei, ea, nphys, nocc, nvir, naux, istart, iend,
                           os_factor=1.0, ss_factor=1.0, vv=None, vev=None):
    """optragf2.c:117-295.  ixq (nocc*nphys, naux), qja (naux, nocc*nvir); accumulates (+=)."""
    ixq = np.asarray(ixq, dtype=np.float64).reshape(nocc, nphys, naux)
    qja = np.asarray(qja, dtype=np.float64).reshape(naux, nocc, nvir)
    vv = np.zeros((nphys, nphys)) if vv is None else vv
    vev = np.zeros((nphys, nphys)) if vev is None else vev
    coul, exch = os_factor + ss_factor, ss_factor
    for i in range(istart, iend):
        xja = (ixq[i] @ qja.reshape(naux, nocc * nvir))                    # :175-188  (x, j*a)
        xia = np.matmul(ixq, qja[:, i, :]).transpose(1, 0, 2).reshape(nphys, nocc * nvir)  # :121-159
        eja = (ei[i] + ei[:, None] - ea[None, :]).ravel()                  # :195-199
        w = coul * xja - exch * xia
        vv += xja @ w.T                                                     # :206-239
        vev += (xja * eja[None, :]) @ w.T                                   # :246-290
    return vv, vev


def np_build_part_loop_uhf(ixq_a, qja_a, qja_b, ei_a, ei_b, ea_a, ea_b, nphys, nocc_a, nocc_b,
                           nvir_a, nvir_b, naux, istart, iend, os_factor=1.0, ss_factor=1.0,
                           vv=None, vev=None):
    """optuagf2.c:125-315 (one spin channel)."""
    ixq_a = np.asarray(ixq_a, dtype=np.float64).reshape(nocc_a, nphys, naux)
    qja_a = np.asarray(qja_a, dtype=np.float64).reshape(naux, nocc_a, nvir_a)
    qja_b = np.asarray(qja_b, dtype=np.float64).reshape(naux, nocc_b * nvir_b)
    vv = np.zeros((nphys, nphys)) if vv is None else vv
    vev = np.zeros((nphys, nphys)) if vev is None else vev
    for i in range(istart, iend):
        xja_aa = ixq_a[i] @ qja_a.reshape(naux, nocc_a * nvir_a)
        xia_aa = np.matmul(ixq_a, qja_a[:, i, :]).transpose(1, 0, 2).reshape(nphys, nocc_a * nvir_a)
        xja_ab = ixq_a[i] @ qja_b
        e_aa = (ei_a[i] + ei_a[:, None] - ea_a[None, :]).ravel()
        e_ab = (ei_a[i] + ei_b[:, None] - ea_b[None, :]).ravel()
        w_aa = ss_factor * (xja_aa - xia_aa)
        vv += xja_aa @ w_aa.T + os_factor * (xja_ab @ xja_ab.T)
        vev += (xja_aa * e_aa[None]) @ w_aa.T + os_factor * ((xja_ab * e_ab[None]) @ xja_ab.T)
    return vv, vev


def np_compress(vv, vev):
    """optragf2.py:270-278: poles (e, c) with  c c^T = vv,  c diag(e) c^T = vev."""
    nphys = vv.shape[0]
    b = np.linalg.cholesky(vv).T
    b_inv = np.linalg.inv(b)
    m = b_inv.T @ vev @ b_inv
    e, c = np.linalg.eigh(m)
    c = b.T @ c[:nphys]
    return e, c


def np_unpack_tril(packed, nphys):
    """pyscf.lib.unpack_tril semantics for a (naux, nphys(nphys+1)/2) array -> (naux, n, n)."""
    naux = packed.shape[0]
    out = np.zeros((naux, nphys, nphys))
    il = np.tril_indices(nphys)
    out[:, il[0], il[1]] = packed
    out[:, il[1], il[0]] = packed
    return out


def np_ao2mo(eri, ci, cj, sym_in="s2"):
    """optragf2.py:175-214 with sym_out='s1': out[Q, i*nj+j] = sum_pq L[Q,p,q] ci[p,i] cj[q,j]."""
    nphys = ci.shape[0]
    l3 = np_unpack_tril(eri, nphys) if sym_in == "s2" else eri.reshape(-1, nphys, nphys)
    half = np.einsum("Qpq,pi->Qiq", l3, ci)
    return np.einsum("Qiq,qj->Qij", half, cj).reshape(l3.shape[0], -1)


# --------------------------------------------------------------------------------------------
# ctypes bindings of the two compiled checkers
# --------------------------------------------------------------------------------------------
_c_lib = None
_ref_lib = None


def build_all(quiet=True):
    """Compile oracle/_build (always) and oracle/_ref (when /root/reference is present)."""
    out = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


def load_c_oracle():
    global _c_lib
    if _c_lib is None:
        path = os.path.join(_HERE, "_build", "liboracle_agf2.so")
        if not os.path.exists(path):
            build_all()
        lib = ctypes.CDLL(path)
        lib.oracle_build_part_loop_rhf.restype = ctypes.c_int
        lib.oracle_build_part_loop_rhf.argtypes = [_f64p, _f64p, _f64p, _f64p] + [_i64] * 6 + \
            [_dbl, _dbl, _f64pw, _f64pw]
        lib.oracle_build_part_loop_uhf.restype = ctypes.c_int
        lib.oracle_build_part_loop_uhf.argtypes = [_f64p] * 7 + [_i64] * 8 + [_dbl, _dbl, _f64pw, _f64pw]
        _c_lib = lib
    return _c_lib


def have_ref():
    return os.path.exists(os.path.join(_HERE, "_ref", "agf2.so"))


def load_ref():
    """The reference's own C loop (oracle/_ref/agf2.so); argtypes as auxgf/lib/agf2.py:28-60."""
    global _ref_lib
    if _ref_lib is None:
        lib = ctypes.CDLL(os.path.join(_HERE, "_ref", "agf2.so"))
        lib.build_part_loop_rhf.restype = None
        lib.build_part_loop_rhf.argtypes = [_f64p, _f64p, _f64p, _f64p] + [_u32] * 6 + [_f64pw, _f64pw]
        lib.build_part_loop_uhf.restype = None
        lib.build_part_loop_uhf.argtypes = [_f64p] * 7 + [_u32] * 8 + [_f64pw, _f64pw]
        _ref_lib = lib
        ref_set_threads(blas=1)      # BLAS is called inside the reference's OpenMP region
    return _ref_lib


def ref_set_threads(omp=None, blas=None):
    """Thread layout of the reference binary: OpenMP team size (the `i` loop, optragf2.c:89,118)
    and the thread count of the OpenBLAS it is linked to (SciPy's bundled one)."""
    import glob
    import scipy
    if blas is not None:
        path = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs",
                                      "libscipy_openblas*.so"))[0]
        ctypes.CDLL(path).scipy_openblas_set_num_threads(int(blas))
    if omp is not None:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(omp))


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def c_build_part_loop_rhf(ixq, qja, ei, ea, nphys, nocc, nvir, naux, istart, iend,
                          os_factor=1.0, ss_factor=1.0):
    vv = np.zeros((nphys, nphys)); vev = np.zeros((nphys, nphys))
    rc = load_c_oracle().oracle_build_part_loop_rhf(
        _c(ixq).reshape(nocc * nphys, naux), _c(qja).reshape(naux, nocc * nvir), _c(ei), _c(ea),
        nphys, nocc, nvir, naux, istart, iend, os_factor, ss_factor, vv, vev)
    assert rc == 0
    return vv, vev


def c_build_part_loop_uhf(ixq_a, qja_a, qja_b, ei_a, ei_b, ea_a, ea_b, nphys, nocc_a, nocc_b,
                          nvir_a, nvir_b, naux, istart, iend, os_factor=1.0, ss_factor=1.0):
    vv = np.zeros((nphys, nphys)); vev = np.zeros((nphys, nphys))
    rc = load_c_oracle().oracle_build_part_loop_uhf(
        _c(ixq_a), _c(qja_a), _c(qja_b), _c(ei_a), _c(ei_b), _c(ea_a), _c(ea_b), nphys, nocc_a,
        nocc_b, nvir_a, nvir_b, naux, istart, iend, os_factor, ss_factor, vv, vev)
    assert rc == 0
    return vv, vev


def ref_build_part_loop_rhf(ixq, qja, ei, ea, nphys, nocc, nvir, naux, istart, iend,
                            vv=None, vev=None):
    vv = np.zeros((nphys, nphys)) if vv is None else vv
    vev = np.zeros((nphys, nphys)) if vev is None else vev
    load_ref().build_part_loop_rhf(_c(ixq).reshape(nocc * nphys, naux),
                                   _c(qja).reshape(naux, nocc * nvir), _c(ei), _c(ea),
                                   nphys, nocc, nvir, naux, istart, iend, vv, vev)
    return vv, vev


def ref_build_part_loop_uhf(ixq_a, qja_a, qja_b, ei_a, ei_b, ea_a, ea_b, nphys, nocc_a, nocc_b,
                            nvir_a, nvir_b, naux, istart, iend, vv=None, vev=None):
    vv = np.zeros((nphys, nphys)) if vv is None else vv
    vev = np.zeros((nphys, nphys)) if vev is None else vev
    load_ref().build_part_loop_uhf(_c(ixq_a), _c(qja_a), _c(qja_b), _c(ei_a), _c(ei_b), _c(ea_a),
                                   _c(ea_b), nphys, nocc_a, nocc_b, nvir_a, nvir_b, naux,
                                   istart, iend, vv, vev)
    return vv, vev


# --------------------------------------------------------------------------------------------
# Seeded synthetic inputs -- duplicated on purpose so the oracle does not
# import the product package.
# --------------------------------------------------------------------------------------------
def synth_rhf(nocc, nvir, naux, nphys, seed=0):
    rng = np.random.default_rng(seed)
    ixq = rng.standard_normal((nocc * nphys, naux)) / np.sqrt(naux)
    qja = rng.standard_normal((naux, nocc * nvir)) / np.sqrt(naux)
    ei = np.sort(rng.uniform(-2.0, -0.5, nocc))
    ea = np.sort(rng.uniform(0.5, 3.0, nvir))
    return ixq, qja, ei, ea


def synth_uhf(nocc_a, nvir_a, nocc_b, nvir_b, naux, nphys, seed=0):
    rng = np.random.default_rng(seed)
    ixq_a = rng.standard_normal((nocc_a * nphys, naux)) / np.sqrt(naux)
    qja_a = rng.standard_normal((naux, nocc_a * nvir_a)) / np.sqrt(naux)
    qja_b = rng.standard_normal((naux, nocc_b * nvir_b)) / np.sqrt(naux)
    ei_a = np.sort(rng.uniform(-2.0, -0.5, nocc_a)); ei_b = np.sort(rng.uniform(-2.0, -0.5, nocc_b))
    ea_a = np.sort(rng.uniform(0.5, 3.0, nvir_a)); ea_b = np.sort(rng.uniform(0.5, 3.0, nvir_b))
    return ixq_a, qja_a, qja_b, ei_a, ei_b, ea_a, ea_b
