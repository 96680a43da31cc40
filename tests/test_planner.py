"""CPU-only: invariants of the work planner (auxgf_b200/csrc/planner.h) through agf2_b200_plan_stats -- the same
code that generates the k1_form tile lists on the GPU, run here without one.  Replaces, as a check, the reference's
`for i in [istart, iend)` loop (auxgf/lib/libagf2/optragf2.c:118) and its MPI split (auxgf/mpi/agf2/optragf2.py:370-374):
every (xi|ja) block that the algorithm needs is formed exactly once over all parts, whatever the sharding."""
import numpy as np
import pytest

from auxgf_b200.lib import agf2 as lib


def _cover(nphys, nocc, nvir, naux, nparts=1, **kw):
    nb = kw.get("nocc_b", 0) if kw.get("uhf") else 0
    cover = np.zeros((nocc, nocc + nb), dtype=np.int64)
    stats = [lib.plan_stats(nphys, nocc, nvir, naux, part=p, nparts=nparts, cover=cover, **kw) for p in range(nparts)]
    return cover, stats


@pytest.mark.parametrize("nparts", [1, 2, 3, 8])
def test_every_pair_block_formed_once(nparts):
    """Full pole range, factorised route: X_ij for all i != j exactly once (two per pole pair), no diagonal blocks."""
    nphys, o, v, naux = 300, 23, 100, 4000                      # o*v >> naux is not needed: force the route instead
    lib.set_coulomb_factorization(1)
    try:
        cover, stats = _cover(nphys, o, v, naux, nparts)
    finally:
        lib.set_coulomb_factorization(2)
    full = stats[0]["blocks_per_pair"]
    assert full == 3 * 13                                        # 3 x-blocks x ceil(100 / 8) column blocks
    expect = full * (1 - np.eye(o, dtype=np.int64))
    assert np.array_equal(cover, expect)
    assert all(s["factorised"] == 1 and s["max_cols"] <= s["cap_cols"] for s in stats)
    ntiles = [s["tiles"] for s in stats]
    assert max(ntiles) - min(ntiles) <= 2 * 3 * 2               # round-robin dealing: parts differ by at most one pair


def test_all_pairs_route_and_slices():
    """Factorisation off: DIAG items appear; a slice [istart, iend) covers X_ij for i in the slice and all j, plus the
    X_ji it needs for the exchange term (j outside the slice)."""
    nphys, o, v, naux = 140, 9, 30, 60
    lib.set_coulomb_factorization(0)
    try:
        cover, stats = _cover(nphys, o, v, naux)
        full = stats[0]["blocks_per_pair"]
        assert stats[0]["factorised"] == 0 and np.array_equal(cover, full * np.ones((o, o), dtype=np.int64))
        cover, stats = _cover(nphys, o, v, naux, nparts=2, istart=2, iend=6)
        inside = np.zeros((o, o), dtype=bool); inside[2:6, 2:6] = True
        rows = np.zeros((o, o), dtype=bool); rows[2:6, :] = True
        cols = np.zeros((o, o), dtype=bool); cols[:, 2:6] = True
        assert np.array_equal(cover[inside], full * np.ones(16, dtype=np.int64))
        assert np.array_equal(cover[rows & ~inside], full * np.ones(4 * 5, dtype=np.int64))      # X_ij, j outside
        assert np.array_equal(cover[cols & ~inside], full * np.ones(4 * 5, dtype=np.int64))      # X_ji for the exchange
        assert not cover[~rows & ~cols].any()
    finally:
        lib.set_coulomb_factorization(2)


def test_uhf_plan():
    """UHF channel: same-spin pairs always; opposite-spin blocks only on the all-pairs route."""
    nphys, oa, va, ob, vb, naux = 92, 8, 84, 7, 85, 300
    for mode, os_blocks in ((1, 0), (0, 1)):
        lib.set_coulomb_factorization(mode)
        try:
            cover, stats = _cover(nphys, oa, va, naux, nparts=3, nocc_b=ob, nvir_b=vb, uhf=True)
        finally:
            lib.set_coulomb_factorization(2)
        full = stats[0]["blocks_per_pair"]                     # 1 x-block x ceil(84 / 8)
        assert np.array_equal(cover[:, :oa], full * (1 - np.eye(oa, dtype=np.int64)))
        assert np.array_equal(cover[:, oa:], os_blocks * 11 * np.ones((oa, ob), dtype=np.int64))  # ceil(85 / 8) = 11


def test_automatic_route_and_thin_tiles():
    """The flop estimate picks the factorisation for the S sectors; nphys = 1100 gives one thin tile row (76 valid rows),
    and the chunk size is tuned to whole waves: 2 waves of 148 CTAs for the A-only pair blocks of the virtual sector."""
    s_vir = lib.plan_stats(1100, 1000, 100, 4000, iend=60)       # a slice keeps the item list short
    assert s_vir["thin_tiles"] > 0 and s_vir["max_cols"] <= s_vir["cap_cols"]
    s_occ = lib.plan_stats(1100, 100, 1000, 4000)
    assert s_occ["factorised"] == 1 and s_occ["thin_tiles"] > 0
    full = lib.plan_stats(1100, 200, 100, 4000)
    assert full["factorised"] == 0                               # o*v = 2e4 is not >> naux: all-pairs route
    small = lib.plan_stats(264, 21, 243, 680)
    assert small["tiles"] > 0 and small["thin_tiles"] > 0        # 264 = 2 x 128 + 8: one 8-row thin block


def test_plan_errors():
    with pytest.raises(lib.Agf2B200Error):
        lib.plan_stats(10, 4, 5, 20, os_factor=-1.0)
    with pytest.raises(lib.Agf2B200Error):
        lib.plan_stats(10, 4, 5, 20, part=2, nparts=2)
    with pytest.raises(lib.Agf2B200Error):
        lib.plan_stats(10, 4, 5, 20, istart=3, iend=9)


def test_coverage_property_random_shapes():
    """Randomised: for any shape, slice, sharding, SCS factors and route, the union over parts of the k1_form tiles forms
    exactly the (xi|ja) blocks the algorithm needs -- X_ij and X_ji for every pole pair with at least one member in the
    slice (the exchange term), X_ii only on the all-pairs route (Coulomb through the pair loop) -- each exactly once."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(o=st.integers(1, 11), v=st.integers(1, 70), p=st.integers(1, 300), naux=st.integers(1, 200),
           nparts=st.integers(1, 5), mode=st.sampled_from([0, 1]), os_f=st.sampled_from([0.0, 1.0, 1.2]),
           ss_f=st.sampled_from([0.0, 1.0, 1.0 / 3.0]), data=st.data())
    def check(o, v, p, naux, nparts, mode, os_f, ss_f, data):
        i0 = data.draw(st.integers(0, o))
        i1 = data.draw(st.integers(i0, o))
        lib.set_coulomb_factorization(mode)
        try:
            cover = np.zeros((o, o), dtype=np.int64)
            stats = [lib.plan_stats(p, o, v, naux, istart=i0, iend=i1, os_factor=os_f, ss_factor=ss_f, part=k,
                                    nparts=nparts, cover=cover) for k in range(nparts)]
        finally:
            lib.set_coulomb_factorization(2)
        full = stats[0]["blocks_per_pair"]
        assert full == ((p + 127) // 128) * ((v + 7) // 8)
        fact = stats[0]["factorised"]
        inside = np.zeros(o, dtype=bool); inside[i0:i1] = True
        expect = np.zeros((o, o), dtype=np.int64)
        for i in range(o):
            for j in range(o):
                if i == j:
                    need = inside[i] and not fact and os_f != 0.0                 # DIAG item
                elif inside[i] and inside[j]:
                    need = ss_f != 0.0 or (not fact and os_f != 0.0)              # PAIR item (A, and S off-route)
                elif inside[i] or inside[j]:
                    need = not (fact and ss_f == 0.0)                             # ASYM item
                else:
                    need = False
                expect[i, j] = full if need else 0
        assert np.array_equal(cover, expect), (o, v, p, i0, i1, nparts, mode, os_f, ss_f)
        assert all(s["max_cols"] <= s["cap_cols"] for s in stats)

    check()


@pytest.mark.parametrize("nocc,nranks,sb", [(10, 2, 3), (17, 3, 2), (64, 8, 1), (250, 8, 9), (33, 4, 100), (8, 8, 5), (41, 5, 4)])
def test_pole_sharded_schedule_covers_every_pair_once(nocc, nranks, sb):
    """The pole-sharded schedule (own block, then visiting sub-blocks split by the parity of i + j) forms every
    unordered pole pair exactly once over all ranks, balanced, and every rank sends its block to every other rank."""
    cover, stats = lib.shard_cover(nocc, nranks, sb)
    assert np.array_equal(cover, np.tril(np.ones((nocc, nocc), dtype=np.int64), -1))
    assert stats[:, 0].sum() == nocc * (nocc - 1) // 2
    for r in range(nranks):
        lo, hi = lib.pole_block(nocc, r, nranks)
        assert stats[r, 1] == (nranks - 1) * (hi - lo)
    if nocc >= 8 * nranks:
        assert stats[:, 0].max() <= 1.15 * stats[:, 0].min()


def test_pole_sharded_sample_is_a_subset_of_the_schedule():
    """shard_step_limit (the benchmark sample of a pole-sharded build): the sampled schedule forms a subset of the pairs,
    each at most once, every rank still talks to every other rank, and limit >= the number of sub-blocks is the full one."""
    nocc, nranks, sb = 97, 4, 5
    full, st_full = lib.shard_cover(nocc, nranks, sb)
    part, st_part = lib.shard_cover(nocc, nranks, sb, limit=2)
    assert part.max() == 1 and np.all(part <= full)
    assert 0 < part.sum() < full.sum()
    assert np.all(st_part[:, 1] > 0) and np.all(st_part[:, 1] < st_full[:, 1])
    same, _ = lib.shard_cover(nocc, nranks, sb, limit=100)
    assert np.array_equal(same, full)
