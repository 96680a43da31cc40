"""CPU-only: the two algebraic identities the GPU path is built on, checked in NumPy against the
oracle's direct statement of the reference loop (auxgf/lib/libagf2/optragf2.c:117-295, optuagf2.c:125-315):

  (i)  pair symmetry          sum_{i>j} A_ij A_ij^T = C - Xc,  A_ij = X_ij - X_ji
  (ii) Coulomb factorisation  sum_{ja} w_j X_ij X_ij^T E_ija^n = ixq_i (e_i^n M0 + n M1) ixq_i^T

so that  vv = ss sum_{i>j in slice} A A^T - ss sum_{i in slice, j outside} X_ij X_ji^T + sum_{i in slice} ixq_i M0 ixq_i^T
with the weights  w_j = os + ss [j outside the slice]  (RHF) -- exactly what planner.h / run_coulomb implement."""
import numpy as np
import pytest

from oracle import agf2_oracle as orc


def _factorised_rhf(ixq, qja, ei, ea, p, o, v, n, i0, i1, os_f, ss_f):
    ix = ixq.reshape(o, p, n)
    q = qja.reshape(n, o, v)
    X = np.einsum("ixQ,Qja->ijxa", ix, q)                       # X[i, j] = (x i | j a)
    E = ei[:, None, None] + ei[None, :, None] - ea[None, None, :]
    vv = np.zeros((p, p)); vev = np.zeros((p, p))
    inside = np.zeros(o, dtype=bool); inside[i0:i1] = True
    for i in range(i0, i1):                                      # exchange-type part: pairs inside the slice ...
        for j in range(i0, i):
            A = X[i, j] - X[j, i]
            vv += ss_f * A @ A.T
            vev += ss_f * (A * E[i, j]) @ A.T
        for j in np.nonzero(~inside)[0]:                         # ... and general products for j outside
            vv -= ss_f * X[i, j] @ X[j, i].T
            vev -= ss_f * (X[i, j] * E[i, j]) @ X[j, i].T
    w = os_f + ss_f * (~inside)                                  # Coulomb-type part through M0 / M1
    qw = q * w[None, :, None]
    M0 = np.einsum("Qja,Rja->QR", qw, q)
    M1 = np.einsum("Qja,ja,Rja->QR", qw, ei[:, None] - ea[None, :], q)
    for i in range(i0, i1):
        vv += ix[i] @ M0 @ ix[i].T
        vev += ix[i] @ (ei[i] * M0 + M1) @ ix[i].T
    return vv, vev


@pytest.mark.parametrize("case", [
    (5, 8, 40, 13, 0, 5, 1.0, 1.0), (6, 7, 30, 11, 2, 5, 1.0, 1.0), (6, 7, 30, 11, 0, 6, 1.2, 1.0 / 3.0),
    (6, 7, 30, 11, 1, 4, 1.2, 1.0 / 3.0), (4, 9, 20, 10, 0, 4, 0.0, 1.0), (4, 9, 20, 10, 0, 4, 1.0, 0.0),
    (7, 3, 12, 9, 3, 4, 1.0, 1.0),
])
def test_rhf_factorised_form_equals_reference_loop(case):
    o, v, n, p, i0, i1, os_f, ss_f = case
    ixq, qja, ei, ea = orc.synth_rhf(o, v, n, p, seed=17)
    ref = orc.np_build_part_loop_rhf(ixq, qja, ei, ea, p, o, v, n, i0, i1, os_f, ss_f)
    got = _factorised_rhf(ixq, qja, ei, ea, p, o, v, n, i0, i1, os_f, ss_f)
    for g, r in zip(got, ref):
        assert np.abs(g - r).max() <= 1e-12 * np.abs(r).max()


def test_uhf_opposite_spin_term_needs_no_integrals():
    """UHF channel, full range: vv = ss sum_{i>j} A A^T (same spin) + os sum_i ixq_i M0^beta ixq_i^T."""
    oa, va, ob, vb, n, p = 5, 9, 4, 10, 30, 12
    ixq_a, qja_a, qja_b, ei_a, ei_b, ea_a, ea_b = orc.synth_uhf(oa, va, ob, vb, n, p, seed=3)
    for os_f, ss_f in ((1.0, 1.0), (1.2, 1.0 / 3.0)):
        ref = orc.np_build_part_loop_uhf(ixq_a, qja_a, qja_b, ei_a, ei_b, ea_a, ea_b, p, oa, ob, va, vb, n, 0, oa, os_f, ss_f)
        ix = ixq_a.reshape(oa, p, n); qa = qja_a.reshape(n, oa, va); qb = qja_b.reshape(n, ob, vb)
        X = np.einsum("ixQ,Qja->ijxa", ix, qa)
        E = ei_a[:, None, None] + ei_a[None, :, None] - ea_a[None, None, :]
        vv = np.zeros((p, p)); vev = np.zeros((p, p))
        for i in range(oa):
            for j in range(i):
                A = X[i, j] - X[j, i]
                vv += ss_f * A @ A.T
                vev += ss_f * (A * E[i, j]) @ A.T
        M0 = os_f * np.einsum("Qja,Rja->QR", qb, qb)
        M1 = os_f * np.einsum("Qja,ja,Rja->QR", qb, ei_b[:, None] - ea_b[None, :], qb)
        for i in range(oa):
            vv += ix[i] @ M0 @ ix[i].T
            vev += ix[i] @ (ei_a[i] * M0 + M1) @ ix[i].T
        assert np.abs(vv - ref[0]).max() <= 1e-12 * np.abs(ref[0]).max()
        assert np.abs(vev - ref[1]).max() <= 1e-12 * np.abs(ref[1]).max()
