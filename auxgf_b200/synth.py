"""Seeded synthetic DF inputs for the moment build (shapes of SURVEY.md section 8d)."""
import numpy as np


def synth_rhf(nocc, nvir, naux, nphys, seed=0):
    """ixq (nocc*nphys, naux), qja (naux, nocc*nvir) ~ N(0,1)/sqrt(naux); sorted pole energies with
    e_occ in U(-2,-0.5), e_vir in U(0.5,3) so that vv is SPD when nocc^2*nvir >= nphys."""
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


def f_alg(o, v, naux, nphys):
    """Algorithmic flops of one sector call: every (xi|ja) formed once,
    two moment contractions, no credit for output symmetry."""
    return 2.0 * nphys * o * o * v * naux + 4.0 * nphys * nphys * o * o * v


def f_alg_uhf(oa, va, ob, vb, naux, nphys):
    return 2.0 * nphys * oa * naux * (oa * va + ob * vb) + 4.0 * nphys * nphys * oa * (oa * va + ob * vb)
