"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz from the reference's own C loop.

Run in the build container (where /root/reference exists):  python oracle/gen_golden.py
Each fixture stores the seed + shape (inputs are regenerated from the seed by
oracle.agf2_oracle.synth_rhf/synth_uhf, which is deterministic for a given NumPy major version --
an input checksum is stored to detect drift) and the (vv, vev) outputs produced by
oracle/_ref/agf2.so = /root/reference/auxgf/lib/libagf2/optragf2.c + optuagf2.c, unmodified.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import agf2_oracle as orc  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# (nocc, nvir, naux, nphys, seed, istart, iend)   -- includes non-tile-multiple and odd sizes
RHF_CASES = [
    (5, 8, 40, 13, 0, 0, 5),
    (5, 8, 40, 13, 1, 1, 4),          # partial [istart, iend) slice (MPI-style)
    (3, 19, 37, 24, 2, 0, 3),         # odd naux / nvir
    (12, 30, 100, 42, 0, 0, 12),
    (7, 100, 130, 50, 3, 0, 7),       # nvir = 100 (64 + 36 split)
    (24, 40, 200, 64, 0, 0, 24),
    (1, 1, 1, 1, 0, 0, 1),            # degenerate
    (6, 5, 16, 130, 4, 0, 6),         # nphys > one 128-row tile
    (5, 70, 50, 77, 5, 0, 5),         # ragged single x-block (10 row blocks) with wide column blocks
    (6, 130, 90, 172, 6, 1, 5),       # full x-block + 44-row remainder, three column blocks, slice
    (9, 64, 300, 140, 7, 0, 9),       # 12-row remainder
    (5, 128, 90, 111, 8, 0, 5),       # ragged single x-block of 14 row blocks (quad tiles), two 64-column blocks
    (6, 72, 60, 223, 9, 1, 6),        # full x-block + 95-row remainder (12 row blocks: quad tiles), 64 + 8 columns, slice
]
# (nocc_a, nvir_a, nocc_b, nvir_b, naux, nphys, seed, istart, iend)
UHF_CASES = [
    (4, 9, 3, 10, 40, 13, 0, 0, 4),
    (8, 20, 7, 21, 60, 28, 1, 0, 8),
    (8, 20, 7, 21, 60, 28, 2, 2, 5),
    (5, 33, 6, 17, 75, 40, 3, 0, 5),
    (5, 66, 4, 71, 60, 150, 4, 0, 5),  # nphys = 128 + 22
    (5, 64, 4, 70, 60, 105, 5, 0, 5),  # nphys = 105: 14 row blocks (quad tiles)
]


def main():
    orc.build_all()
    assert orc.have_ref(), "oracle/_ref/agf2.so missing (needs /root/reference)"
    os.makedirs(OUT, exist_ok=True)
    force = "--force" in sys.argv      # existing fixtures are kept byte-identical unless asked otherwise
    wrote = 0
    for k, (o, v, n, p, seed, i0, i1) in enumerate(RHF_CASES):
        if os.path.exists(os.path.join(OUT, "rhf_%02d.npz" % k)) and not force:
            continue
        wrote += 1
        ixq, qja, ei, ea = orc.synth_rhf(o, v, n, p, seed)
        vv, vev = orc.ref_build_part_loop_rhf(ixq, qja, ei, ea, p, o, v, n, i0, i1)
        np.savez(os.path.join(OUT, "rhf_%02d.npz" % k), shape=np.array([o, v, n, p, seed, i0, i1]),
                 in_sum=np.array([ixq.sum(), qja.sum(), ei.sum(), ea.sum()]), vv=vv, vev=vev)
    for k, (oa, va, ob, vb, n, p, seed, i0, i1) in enumerate(UHF_CASES):
        if os.path.exists(os.path.join(OUT, "uhf_%02d.npz" % k)) and not force:
            continue
        wrote += 1
        args = orc.synth_uhf(oa, va, ob, vb, n, p, seed)
        vv, vev = orc.ref_build_part_loop_uhf(*args, p, oa, ob, va, vb, n, i0, i1)
        np.savez(os.path.join(OUT, "uhf_%02d.npz" % k),
                 shape=np.array([oa, va, ob, vb, n, p, seed, i0, i1]),
                 in_sum=np.array([a.sum() for a in args]), vv=vv, vev=vev)
    print("wrote", wrote, "new fixtures to", OUT, "(%d cases in total)" % (len(RHF_CASES) + len(UHF_CASES)))


if __name__ == "__main__":
    main()
