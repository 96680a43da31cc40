"""Energies from poles (``auxgf/aux/energy.py:9-111``): Galitskii-Migdal two-body energy and its MP2
special case.  ``energy_2body_aux`` is restated as one matrix product (the reference loops over the
occupied Green's-function poles with a five-operand einsum, energy.py:43-49)."""
import numpy as np

from auxgf_b200 import backend


def energy_2body_aux(gf, se, both_sides=False):
    if isinstance(se, (tuple, list)):
        n = len(se)
        gfs = gf if isinstance(gf, (tuple, list)) else (gf,) * n
        return sum(energy_2body_aux(gfs[i], se[i], both_sides=both_sides) for i in range(n)) / n
    nphys = se.nphys

    be = backend.get()

    def half(gv, ge, sv, sen, sign):
        # sum_{l,k} (sum_x gv[x,l] sv[x,k])^2 * sign / (ge[l] - sen[k]): one GEMM + reduction, on the device
        return sign * be.energy_2body(gv[:nphys], ge, sv, sen)

    go, so = gf.e < gf.chempot, se.e < se.chempot
    e2b = half(gf.v[:, go], gf.e[go], se.v[:, ~so], se.e[~so], 1.0)
    if both_sides:
        e2b += half(gf.v[:, ~go], gf.e[~go], se.v[:, so], se.e[so], -1.0)
    else:
        e2b *= 2.0
    return e2b


def energy_mp2_aux(mo, se, both_sides=False):
    if isinstance(se, (tuple, list)):
        n = len(se)
        mos = mo if np.ndim(mo) == 2 or isinstance(mo, (tuple, list)) else (mo,) * n
        return sum(energy_mp2_aux(np.asarray(mos[i]), se[i], both_sides=both_sides) for i in range(n)) / n
    mo = np.asarray(mo, dtype=np.float64)
    occ = mo < se.chempot
    vxk = se.v_vir[occ]
    e2b = float(np.sum(vxk * vxk / (mo[occ][:, None] - se.e_vir[None, :])))
    if both_sides:
        vxk = se.v_occ[~occ]
        e2b += float(np.sum(vxk * vxk * (-1.0 / (mo[~occ][:, None] - se.e_occ[None, :]))))
        e2b *= 0.5
    return e2b
