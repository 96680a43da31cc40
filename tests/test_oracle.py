"""CPU-only: pins the oracle (oracle/agf2_oracle.{c,py}) against the golden vectors produced by the
reference's own C loop (oracle/_ref/agf2.so, built from /root/reference by oracle/Makefile) and, when
that binary is present, against the binary itself."""
import glob
import os

import numpy as np
import pytest

from oracle import agf2_oracle as orc

TOL = 1e-12   # relative to max|ref|; FP64 summation-order noise is ~1e-15


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def _rhf_cases(golden_dir):
    return sorted(glob.glob(os.path.join(golden_dir, "rhf_*.npz")))


def _uhf_cases(golden_dir):
    return sorted(glob.glob(os.path.join(golden_dir, "uhf_*.npz")))


def test_golden_present(golden_dir):
    assert len(_rhf_cases(golden_dir)) >= 8 and len(_uhf_cases(golden_dir)) >= 4


def test_rhf_oracles_match_golden(golden_dir):
    for path in _rhf_cases(golden_dir):
        g = np.load(path)
        o, v, n, p, seed, i0, i1 = [int(x) for x in g["shape"]]
        ixq, qja, ei, ea = orc.synth_rhf(o, v, n, p, seed)
        np.testing.assert_allclose([ixq.sum(), qja.sum(), ei.sum(), ea.sum()], g["in_sum"], rtol=1e-13)
        for fn in (orc.np_build_part_loop_rhf, orc.c_build_part_loop_rhf):
            vv, vev = fn(ixq, qja, ei, ea, p, o, v, n, i0, i1)
            assert relerr(vv, g["vv"]) < TOL and relerr(vev, g["vev"]) < TOL, (path, fn.__name__)


def test_uhf_oracles_match_golden(golden_dir):
    for path in _uhf_cases(golden_dir):
        g = np.load(path)
        oa, va, ob, vb, n, p, seed, i0, i1 = [int(x) for x in g["shape"]]
        args = orc.synth_uhf(oa, va, ob, vb, n, p, seed)
        for fn in (orc.np_build_part_loop_uhf, orc.c_build_part_loop_uhf):
            vv, vev = fn(*args, p, oa, ob, va, vb, n, i0, i1)
            assert relerr(vv, g["vv"]) < TOL and relerr(vev, g["vev"]) < TOL, (path, fn.__name__)


@pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref/agf2.so not built")
def test_reference_binary_matches_golden_and_oracle(golden_dir):
    g = np.load(_rhf_cases(golden_dir)[3])
    o, v, n, p, seed, i0, i1 = [int(x) for x in g["shape"]]
    ixq, qja, ei, ea = orc.synth_rhf(o, v, n, p, seed)
    vv, vev = orc.ref_build_part_loop_rhf(ixq, qja, ei, ea, p, o, v, n, i0, i1)
    assert relerr(vv, g["vv"]) < TOL and relerr(vev, g["vev"]) < TOL
    # a fresh, non-fixture shape: reference binary vs NumPy restatement
    ixq, qja, ei, ea = orc.synth_rhf(9, 21, 55, 33, 7)
    r = orc.ref_build_part_loop_rhf(ixq, qja, ei, ea, 33, 9, 21, 55, 2, 8)
    m = orc.np_build_part_loop_rhf(ixq, qja, ei, ea, 33, 9, 21, 55, 2, 8)
    assert relerr(m[0], r[0]) < TOL and relerr(m[1], r[1]) < TOL
    a = orc.synth_uhf(6, 11, 5, 12, 44, 17, 5)
    r = orc.ref_build_part_loop_uhf(*a, 17, 6, 5, 11, 12, 44, 1, 6)
    m = orc.np_build_part_loop_uhf(*a, 17, 6, 5, 11, 12, 44, 1, 6)
    assert relerr(m[0], r[0]) < TOL and relerr(m[1], r[1]) < TOL


def test_slices_sum_to_full_and_full_is_symmetric():
    ixq, qja, ei, ea = orc.synth_rhf(6, 9, 30, 14, 11)
    full = orc.np_build_part_loop_rhf(ixq, qja, ei, ea, 14, 6, 9, 30, 0, 6)
    parts = [orc.np_build_part_loop_rhf(ixq, qja, ei, ea, 14, 6, 9, 30, a, b) for a, b in ((0, 2), (2, 5), (5, 6))]
    for k in range(2):
        assert relerr(sum(p[k] for p in parts), full[k]) < TOL
        assert relerr(full[k], full[k].T) < TOL
    assert relerr(parts[0][0], parts[0][0].T) > 1e-6   # a partial slice is NOT symmetric (exchange term)


def test_scs_factors_follow_dfrmp2():
    """os/ss enter as Coulomb (os+ss) and exchange (ss): auxgf/aux/dfrmp2.py:82-84,107-109."""
    o, v, n, p = 4, 6, 20, 9
    ixq, qja, ei, ea = orc.synth_rhf(o, v, n, p, 3)
    x = np.einsum("ixq,qja->xija", ixq.reshape(o, p, n), qja.reshape(n, o, v))
    os_f, ss_f = 1.2, 1.0 / 3.0
    # explicit pole list of dfrmp2.py: alpha-beta poles (os) and same-spin antisymmetrised poles (ss)
    vv_ref = os_f * np.einsum("xija,yija->xy", x, x) + \
        ss_f * np.einsum("xija,yija->xy", x, x - x.transpose(0, 2, 1, 3))
    vv, _ = orc.np_build_part_loop_rhf(ixq, qja, ei, ea, p, o, v, n, 0, o, os_f, ss_f)
    assert relerr(vv, vv_ref) < TOL
    vvc, _ = orc.c_build_part_loop_rhf(ixq, qja, ei, ea, p, o, v, n, 0, o, os_f, ss_f)
    assert relerr(vvc, vv_ref) < TOL


def test_compress_reproduces_moments():
    ixq, qja, ei, ea = orc.synth_rhf(6, 9, 30, 14, 2)
    vv, vev = orc.np_build_part_loop_rhf(ixq, qja, ei, ea, 14, 6, 9, 30, 0, 6)
    e, c = orc.np_compress(vv, vev)
    assert relerr(c @ c.T, vv) < 1e-11 and relerr((c * e) @ c.T, vev) < 1e-11
