"""GPU parity tests: the CUDA moment build, called through the C-ABI (auxgf_b200.lib.agf2 ->
libagf2_b200.so), against the oracle (oracle/agf2_oracle.py, pinned in tests/test_oracle.py) on the
same seeded inputs, against the committed golden vectors of the reference's own C loop, and through
size-independent properties.  Tolerance: moments within 1e-10 relative (BASELINE.json north_star)."""
import glob
import os

import numpy as np
import pytest

from oracle import agf2_oracle as orc

pytestmark = pytest.mark.gpu
TOL = 1e-10


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.fixture(scope="module")
def lib():
    from auxgf_b200.lib import agf2 as lib
    return lib


def _check_rhf(lib, o, v, n, p, seed, i0=0, i1=None, os_f=1.0, ss_f=1.0, tol=TOL):
    i1 = o if i1 is None else i1
    ixq, qja, ei, ea = orc.synth_rhf(o, v, n, p, seed)
    vv, vev = lib.moments_rhf(ixq, qja, ei, ea, p, i0, i1, os_f, ss_f)
    rv, rev = orc.np_build_part_loop_rhf(ixq, qja, ei, ea, p, o, v, n, i0, i1, os_f, ss_f)
    assert relerr(vv, rv) < tol and relerr(vev, rev) < tol, (relerr(vv, rv), relerr(vev, rev))
    return vv, vev


@pytest.mark.parametrize("simple", [True, False])
def test_golden_rhf(lib, golden_dir, simple):
    lib.set_simple_kernels(simple)
    try:
        for path in sorted(glob.glob(os.path.join(golden_dir, "rhf_*.npz"))):
            g = np.load(path)
            o, v, n, p, seed, i0, i1 = [int(x) for x in g["shape"]]
            ixq, qja, ei, ea = orc.synth_rhf(o, v, n, p, seed)
            vv, vev = lib.moments_rhf(ixq, qja, ei, ea, p, i0, i1)
            assert relerr(vv, g["vv"]) < TOL and relerr(vev, g["vev"]) < TOL, \
                (path, simple, relerr(vv, g["vv"]), relerr(vev, g["vev"]))
    finally:
        lib.set_simple_kernels(False)


@pytest.mark.parametrize("simple", [True, False])
def test_golden_uhf(lib, golden_dir, simple):
    lib.set_simple_kernels(simple)
    try:
        for path in sorted(glob.glob(os.path.join(golden_dir, "uhf_*.npz"))):
            g = np.load(path)
            oa, va, ob, vb, n, p, seed, i0, i1 = [int(x) for x in g["shape"]]
            args = orc.synth_uhf(oa, va, ob, vb, n, p, seed)
            vv, vev = lib.moments_uhf(*args, p, i0, i1)
            assert relerr(vv, g["vv"]) < TOL and relerr(vev, g["vev"]) < TOL, \
                (path, simple, relerr(vv, g["vv"]), relerr(vev, g["vev"]))
    finally:
        lib.set_simple_kernels(False)


@pytest.mark.parametrize("shape", [
    (5, 8, 40, 13), (3, 19, 37, 24), (12, 30, 100, 42), (7, 100, 130, 50), (6, 5, 16, 130),
    (24, 240, 1000, 264), (40, 100, 333, 141), (9, 65, 129, 257), (2, 7, 15, 3),
    (5, 19, 84, 24),          # water cc-pVDZ scale (BASELINE.json configs[0]: P=24, o=5, v=19)
    (21, 243, 680, 264),      # benzene cc-pVTZ / cc-pVTZ-RI scale, HF shape (configs[1])
])
def test_parity_ladder_rhf(lib, shape):
    o, v, n, p = shape
    _check_rhf(lib, o, v, n, p, seed=1)


def test_thin_tiles(lib):
    """Last x-block with 1..10 valid 8-row blocks (k1_mainloop_thin<1..10>), alone (nphys < 128: it also writes
    the energy weights) and behind a full x-block; wide and narrow column blocks; slice; UHF."""
    for mbv in range(1, 11):
        _check_rhf(lib, 4, 70, 50, 8 * mbv - 3, seed=30 + mbv)
    for p in (128 + 5, 128 + 44, 256 + 77, 128 + 80):
        _check_rhf(lib, 5, 130, 90, p, seed=41)
    _check_rhf(lib, 6, 64, 40, 140, seed=42, i0=1, i1=4, os_f=1.2, ss_f=1.0 / 3.0)
    args = orc.synth_uhf(5, 66, 4, 71, 60, 150, 5)
    vv, vev = lib.moments_uhf(*args, 150)
    rv, rev = orc.np_build_part_loop_uhf(*args, 150, 5, 4, 66, 71, 60, 0, 5)
    assert relerr(vv, rv) < TOL and relerr(vev, rev) < TOL
    lib.set_coulomb_factorization(0)      # DIAG / OS items (single-block and paired plain tiles) through thin tiles
    try:
        _check_rhf(lib, 5, 130, 90, 128 + 44, seed=43)
        vv, vev = lib.moments_uhf(*args, 150)
        assert relerr(vv, rv) < TOL and relerr(vev, rev) < TOL
    finally:
        lib.set_coulomb_factorization(2)


def test_quad_tiles(lib):
    """Last x-block with 11..14 valid 8-row blocks (k1_quad_tile<6>, <7>: 2 x 4 warp grid), alone and behind full
    x-blocks; 64-column blocks (where the mapping is chosen) next to narrower ones; both routes; slice; UHF; and the
    planner really emits such tiles for these shapes."""
    for p_ in (8 * 11 - 5, 8 * 12, 8 * 13 - 1, 8 * 14 - 7, 128 + 8 * 12 - 2, 256 + 111):
        _check_rhf(lib, 5, 128, 90, p_, seed=50 + p_ % 7)
        _check_rhf(lib, 4, 100, 50, p_, seed=60 + p_ % 5)
    st = lib.plan_stats(367, 10, 128, 400)
    assert st["thin_tiles"] > 0
    _check_rhf(lib, 6, 64, 40, 128 + 100, seed=44, i0=1, i1=4, os_f=1.2, ss_f=1.0 / 3.0)
    args = orc.synth_uhf(5, 66, 4, 71, 60, 111, 6)
    rv, rev = orc.np_build_part_loop_uhf(*args, 111, 5, 4, 66, 71, 60, 0, 5)
    for mode in (2, 0):                   # 0: DIAG / OS items (single-block and paired plain tiles) through quad tiles
        lib.set_coulomb_factorization(mode)
        try:
            vv, vev = lib.moments_uhf(*args, 111)
            assert relerr(vv, rv) < TOL and relerr(vev, rev) < TOL
            _check_rhf(lib, 5, 130, 90, 128 + 105, seed=45)
        finally:
            lib.set_coulomb_factorization(2)


def test_reference_drop_in_symbols(lib):
    """The void symbols with the reference's exact signature (optragf2.c:32-43, optuagf2.c:31-47)."""
    o, v, n, p = 6, 21, 50, 30
    ixq, qja, ei, ea = orc.synth_rhf(o, v, n, p, 4)
    vv = np.full((p, p), 0.5); vev = np.full((p, p), -0.25)          # `+=` semantics
    lib._libagf2.build_part_loop_rhf(ixq, qja, ei, ea, p, o, v, n, 1, 5, vv, vev)
    rv, rev = orc.np_build_part_loop_rhf(ixq, qja, ei, ea, p, o, v, n, 1, 5)
    assert relerr(vv - 0.5, rv) < TOL and relerr(vev + 0.25, rev) < TOL
    a = orc.synth_uhf(5, 12, 4, 13, 40, 19, 6)
    vv = np.zeros((19, 19)); vev = np.zeros((19, 19))
    lib._libagf2.build_part_loop_uhf(*[np.ascontiguousarray(x) for x in a], 19, 5, 4, 12, 13, 40, 0, 5, vv, vev)
    rv, rev = orc.np_build_part_loop_uhf(*a, 19, 5, 4, 12, 13, 40, 0, 5)
    assert relerr(vv, rv) < TOL and relerr(vev, rev) < TOL


def test_slices_rhf(lib):
    o, v, n, p = 10, 33, 70, 45
    ixq, qja, ei, ea = orc.synth_rhf(o, v, n, p, 2)
    full = lib.moments_rhf(ixq, qja, ei, ea, p)
    acc = [np.zeros((p, p)), np.zeros((p, p))]
    for a, b in ((0, 3), (3, 4), (4, 10)):
        part = _check_rhf(lib, o, v, n, p, 2, a, b)
        acc[0] += part[0]; acc[1] += part[1]
    assert relerr(acc[0], full[0]) < TOL and relerr(acc[1], full[1]) < TOL
    assert relerr(full[0], full[0].T) < 1e-13     # full range: exactly symmetric by construction


def test_scs_factors(lib):
    _check_rhf(lib, 8, 20, 60, 31, seed=5, os_f=1.2, ss_f=1.0 / 3.0)
    _check_rhf(lib, 8, 20, 60, 31, seed=5, i0=2, i1=6, os_f=1.2, ss_f=1.0 / 3.0)
    _check_rhf(lib, 8, 20, 60, 31, seed=5, os_f=0.0, ss_f=1.0)
    _check_rhf(lib, 8, 20, 60, 31, seed=5, os_f=1.0, ss_f=0.0)


@pytest.mark.parametrize("shape", [(8, 84, 7, 85, 300, 92), (4, 9, 3, 10, 40, 13), (20, 40, 19, 41, 128, 60)])
def test_parity_uhf(lib, shape):
    oa, va, ob, vb, n, p = shape
    args = orc.synth_uhf(oa, va, ob, vb, n, p, 3)
    for (i0, i1, os_f, ss_f) in ((0, oa, 1.0, 1.0), (1, oa - 1, 1.0, 1.0), (0, oa, 1.2, 1.0 / 3.0)):
        vv, vev = lib.moments_uhf(*args, p, i0, i1, os_f, ss_f)
        rv, rev = orc.np_build_part_loop_uhf(*args, p, oa, ob, va, vb, n, i0, i1, os_f, ss_f)
        assert relerr(vv, rv) < TOL and relerr(vev, rev) < TOL, (shape, i0, i1, relerr(vv, rv), relerr(vev, rev))


def test_coulomb_factorization_on_off(lib):
    """The Coulomb factorisation (M0/M1 route) and the all-pairs route give the same moments: RHF, RHF slice,
    SCS, UHF, UHF slice -- each against the oracle with the factorisation on and off."""
    for on in (True, False):
        lib.set_coulomb_factorization(on)
        try:
            _check_rhf(lib, 9, 27, 150, 37, seed=21)
            _check_rhf(lib, 9, 27, 150, 37, seed=21, i0=2, i1=7, os_f=1.2, ss_f=1.0 / 3.0)
            _check_rhf(lib, 3, 140, 70, 150, seed=22)               # naux < one tile, nphys > one tile
            args = orc.synth_uhf(6, 17, 5, 18, 90, 23, 4)
            for (i0, i1, os_f, ss_f) in ((0, 6, 1.0, 1.0), (2, 5, 1.2, 1.0 / 3.0), (0, 6, 1.0, 0.0), (1, 6, 0.0, 1.0)):
                vv, vev = lib.moments_uhf(*args, 23, i0, i1, os_f, ss_f)
                rv, rev = orc.np_build_part_loop_uhf(*args, 23, 6, 5, 17, 18, 90, i0, i1, os_f, ss_f)
                assert relerr(vv, rv) < TOL and relerr(vev, rev) < TOL, (on, i0, i1, relerr(vv, rv), relerr(vev, rev))
        finally:
            lib.set_coulomb_factorization(2)


def test_small_workspace_many_chunks(lib):
    """Force many chunks (tiny workspace) -- exercises the chunk loop and tile double-buffering."""
    lib.set_workspace_bytes(1 << 20)
    try:
        _check_rhf(lib, 12, 30, 100, 140, seed=8)
        _check_rhf(lib, 12, 30, 100, 140, seed=8, i0=3, i1=9)
    finally:
        lib.set_workspace_bytes(40 << 20)


def test_properties_medium(lib):
    """Size-independent properties on a shape the NumPy oracle would need minutes for:
    symmetry, PSD of vv, energy-shift identity  vev(e + d) - vev(e) = d * vv  (E = e_i+e_j-e_a is
    linear in a uniform shift d of the occupied energies: shift of 2d - 0), and linearity in ixq."""
    o, v, n, p = 48, 200, 1024, 400
    ixq, qja, ei, ea = orc.synth_rhf(o, v, n, p, 9)
    vv, vev = lib.moments_rhf(ixq, qja, ei, ea, p)
    assert relerr(vv, vv.T) < 1e-13 and relerr(vev, vev.T) < 1e-13
    assert np.linalg.eigvalsh(vv).min() > 0
    d = 0.37
    vv2, vev2 = lib.moments_rhf(ixq, qja, ei + d, ea - d, p)      # E -> E + 3d
    assert relerr(vv2, vv) < 1e-12
    assert relerr(vev2 - vev, 3 * d * vv) < 1e-9
    vv3, vev3 = lib.moments_rhf(2.0 * ixq, qja, ei, ea, p)          # quadratic in ixq
    assert relerr(vv3, 4 * vv) < 1e-12 and relerr(vev3, 4 * vev) < 1e-12
    # one slice of the same problem against the reference binary / NumPy oracle
    r = orc.np_build_part_loop_rhf(ixq, qja, ei, ea, p, o, v, n, 5, 7)
    g = lib.moments_rhf(ixq, qja, ei, ea, p, 5, 7)
    assert relerr(g[0], r[0]) < TOL and relerr(g[1], r[1]) < TOL


def test_compress_and_eigh(lib):
    o, v, n, p = 12, 30, 100, 42
    ixq, qja, ei, ea = orc.synth_rhf(o, v, n, p, 0)
    vv, vev = lib.moments_rhf(ixq, qja, ei, ea, p)
    e, c = lib.compress(vv, vev)
    assert relerr(c @ c.T, vv) < 1e-10 and relerr((c * e) @ c.T, vev) < 1e-10
    e_ref, _ = orc.np_compress(vv, vev)
    np.testing.assert_allclose(e, e_ref, rtol=0, atol=1e-9)
    a = np.random.default_rng(0).standard_normal((97, 97)); a = a + a.T
    w, u = lib.eigh(a)
    np.testing.assert_allclose(w, np.linalg.eigvalsh(a), rtol=0, atol=1e-11)
    assert relerr((u * w) @ u.T, a) < 1e-12
    with pytest.raises(lib.Agf2B200Error):
        lib.compress(-np.eye(5), np.eye(5))      # not SPD -> status, like the reference's LinAlgError


def test_device_resident_async_and_sharded(lib):
    """Device-resident C-ABI: two back-to-back asynchronous calls (occupied + virtual sector, as bench.py
    issues them) and the part/nparts sharding used for multi-GPU (sum over parts == full build)."""
    import torch
    dev = torch.device("cuda", 0)
    o, v, n, p = 14, 50, 120, 70
    ixq, qja, ei, ea = orc.synth_rhf(o, v, n, p, 12)
    ixq2, qja2, ei2, ea2 = orc.synth_rhf(v, o, n, p, 13)      # virtual-sector call: roles swapped
    T = lambda a, sh: torch.from_numpy(np.ascontiguousarray(a)).reshape(sh).to(dev)  # noqa: E731
    d1 = (T(ixq, (o, p, n)), T(qja, (n, o, v)), T(ei, (o,)), T(ea, (v,)))
    d2 = (T(ixq2, (v, p, n)), T(qja2, (n, v, o)), T(ei2, (v,)), T(ea2, (o,)))
    out = torch.zeros((2, 2, p, p), dtype=torch.float64, device=dev)
    for _ in range(3):      # repeated: scratch / staging reuse across in-flight calls
        out.zero_()
        lib.moments_rhf_dev(*d1, vv=out[0, 0], vev=out[0, 1], sync=False)
        lib.moments_rhf_dev(*d2, vv=out[1, 0], vev=out[1, 1], sync=False)
    torch.cuda.synchronize()
    r1 = orc.np_build_part_loop_rhf(ixq, qja, ei, ea, p, o, v, n, 0, o)
    r2 = orc.np_build_part_loop_rhf(ixq2, qja2, ei2, ea2, p, v, o, n, 0, v)
    h = out.cpu().numpy()
    assert relerr(h[0, 0], r1[0]) < TOL and relerr(h[0, 1], r1[1]) < TOL
    assert relerr(h[1, 0], r2[0]) < TOL and relerr(h[1, 1], r2[1]) < TOL
    # sharding: 3 parts accumulate into the same output
    acc = torch.zeros((2, p, p), dtype=torch.float64, device=dev)
    for part in range(3):
        lib.moments_rhf_dev(*d1, part=part, nparts=3, vv=acc[0], vev=acc[1], sync=True)
    a = acc.cpu().numpy()
    assert relerr(a[0], r1[0]) < TOL and relerr(a[1], r1[1]) < TOL
    # padded leading dimensions (odd naux / nvir -> even pitch)
    o, v, n, p = 5, 9, 21, 17
    ixq, qja, ei, ea = orc.synth_rhf(o, v, n, p, 14)
    ip = torch.zeros((o, p, n + 1), dtype=torch.float64, device=dev); ip[:, :, :n] = T(ixq, (o, p, n))
    qp = torch.zeros((n, o, v + 1), dtype=torch.float64, device=dev); qp[:, :, :v] = T(qja, (n, o, v))
    vv, vev = lib.moments_rhf_dev(ip, qp, T(ei, (o,)), T(ea, (v,)), naux=n)
    r = orc.np_build_part_loop_rhf(ixq, qja, ei, ea, p, o, v, n, 0, o)
    assert relerr(vv.cpu().numpy(), r[0]) < TOL and relerr(vev.cpu().numpy(), r[1]) < TOL


def test_edge_cases_and_errors(lib):
    """Empty slice, single pole / single virtual, naux = 1, sharding into more parts than work items; argument errors
    come back as status codes (Agf2B200Error), never as a crash or a silent CPU path."""
    import torch
    dev = torch.device("cuda", 0)
    o, v, n, p = 6, 9, 20, 14
    ixq, qja, ei, ea = orc.synth_rhf(o, v, n, p, 50)
    vv, vev = lib.moments_rhf(ixq, qja, ei, ea, p, 3, 3)            # empty slice: nothing accumulated
    assert not vv.any() and not vev.any()
    _check_rhf(lib, 1, 1, 1, 1, seed=51)
    _check_rhf(lib, 1, 7, 5, 9, seed=52)                            # one pole: only the Coulomb-type term
    _check_rhf(lib, 7, 1, 5, 9, seed=53)
    _check_rhf(lib, 3, 5, 1, 140, seed=54)                          # naux = 1
    T = lambda a, sh: torch.from_numpy(np.ascontiguousarray(a)).reshape(sh).to(dev)  # noqa: E731
    # even pitch required by the device entry point: pad v = 9 -> 10
    qp = torch.zeros((n, o, v + 1), dtype=torch.float64, device=dev); qp[:, :, :v] = T(qja, (n, o, v))
    d = (T(ixq, (o, p, n)), qp, T(ei, (o,)), T(ea, (v,)))
    acc = torch.zeros((2, p, p), dtype=torch.float64, device=dev)
    for part in range(40):                                          # 40 parts > 21 pole pairs and > naux / 64 blocks
        lib.moments_rhf_dev(*d, part=part, nparts=40, vv=acc[0], vev=acc[1])
    r = orc.np_build_part_loop_rhf(ixq, qja, ei, ea, p, o, v, n, 0, o)
    a = acc.cpu().numpy()
    assert relerr(a[0], r[0]) < TOL and relerr(a[1], r[1]) < TOL
    with pytest.raises(lib.Agf2B200Error):
        lib.moments_rhf(ixq, qja, ei, ea, p, 0, o, -1.0, 1.0)       # negative SCS factor
    vv, vev = lib.moments_rhf(ixq, qja, ei, ea, p, 4, 2)            # istart > iend: the reference's loop runs zero times
    assert not vv.any() and not vev.any()
    with pytest.raises(lib.Agf2B200Error):
        lib.moments_rhf_dev(*d, part=3, nparts=2)
    with pytest.raises(lib.Agf2B200Error):
        lib.moments_rhf_dev(T(ixq, (o, p, n)), T(qja, (n, o, v)), T(ei, (o,)), T(ea, (v,)))   # odd pitch (v = 9)


def test_bit_repeatability(lib):
    """Deterministic mode (default): stream-K partial tiles are summed in CTA order by k2_fixup and every stream has
    its own accumulator set, so the same build gives the same bits -- many chunks on two streams, Coulomb
    factorisation, slices, UHF, host-pointer and device-resident entry points."""
    import torch
    lib.set_workspace_bytes(2 << 20)          # many chunks: both streams busy
    try:
        o, v, n, p = 20, 60, 200, 300
        ixq, qja, ei, ea = orc.synth_rhf(o, v, n, p, 70)
        runs = [lib.moments_rhf(ixq, qja, ei, ea, p) for _ in range(3)]
        for r in runs[1:]:
            assert np.array_equal(r[0], runs[0][0]) and np.array_equal(r[1], runs[0][1])
        a = [lib.moments_rhf(ixq, qja, ei, ea, p, 3, 11, 1.2, 1.0 / 3.0) for _ in range(2)]
        assert np.array_equal(a[0][0], a[1][0]) and np.array_equal(a[0][1], a[1][1])
        args = orc.synth_uhf(12, 40, 11, 41, 150, 200, 71)
        u = [lib.moments_uhf(*args, 200) for _ in range(2)]
        assert np.array_equal(u[0][0], u[1][0]) and np.array_equal(u[0][1], u[1][1])
        dev = torch.device("cuda", 0)
        T = lambda x, sh: torch.from_numpy(np.ascontiguousarray(x)).reshape(sh).to(dev)  # noqa: E731
        d = (T(ixq, (o, p, n)), T(qja, (n, o, v)), T(ei, (o,)), T(ea, (v,)))
        g = [torch.stack(lib.moments_rhf_dev(*d)).cpu().numpy() for _ in range(2)]
        assert np.array_equal(g[0], g[1])
        assert np.array_equal(g[0][0], runs[0][0])          # host-pointer and device-resident paths agree bit for bit
        # the non-deterministic mode (red.global.add) agrees to rounding
        lib.set_deterministic(False)
        nd = lib.moments_rhf(ixq, qja, ei, ea, p)
        assert relerr(nd[0], runs[0][0]) < 1e-13 and relerr(nd[1], runs[0][1]) < 1e-13
    finally:
        lib.set_deterministic(True)
        lib.set_workspace_bytes(40 << 20)


@pytest.mark.parametrize("shape", [(300, 200, 96), (256, 64, 16), (1, 1, 1), (130, 70, 37), (513, 129, 260),
                                   (2000, 1000, 555), (64, 3000, 48)])
def test_general_gemm(lib, shape):
    """agf2_b200_ctx_gemm (k2_moment general-GEMM instantiation, 256 x 64 tiles, stream-K + fixed-order fixup)
    against torch.matmul in FP64: plain, transposed output, accumulate, padded pitches; bit-reproducible."""
    import torch
    m, n, k = shape
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(m * 7 + n)
    kp = k + (k & 1) + 2                      # padded (even) pitch
    a = torch.randn((m, kp), generator=g, device=dev, dtype=torch.float64)[:, :k]
    b = torch.randn((n, kp), generator=g, device=dev, dtype=torch.float64)[:, :k]
    ref = a @ b.T
    scale = float(ref.abs().max()) + 1e-300
    c = lib.gemm_dev(a, b)
    assert float((c - ref).abs().max()) / scale < 1e-13
    assert torch.equal(c, lib.gemm_dev(a, b))
    ct = lib.gemm_dev(a, b, transpose=True)
    assert torch.equal(ct, c.T.contiguous())
    acc = torch.full((m, n + 3), 2.0, device=dev, dtype=torch.float64)[:, :n]     # pitch n + 3, += semantics
    lib.gemm_dev(a, b, out=acc, accumulate=True)
    assert float((acc - 2.0 - ref).abs().max()) / scale < 1e-13


def test_contexts(lib):
    """User-created contexts: own options (the process defaults are untouched), own scratch; destroying one
    leaves the default context working; the void drop-ins surface failures through the sticky status."""
    import ctypes
    h = ctypes.c_void_p()
    lib.check(lib._libagf2.agf2_b200_ctx_create(-1, ctypes.byref(h)))
    lib.set_option("workspace_bytes", 1 << 20, ctx=h)
    lib.set_option("deterministic", 0, ctx=h)
    dev = ctypes.c_int(-1)
    lib.check(lib._libagf2.agf2_b200_ctx_get_device(h, ctypes.byref(dev)))
    assert dev.value == 0
    o, v, n, p = 9, 30, 64, 50
    ixq, qja, ei, ea = orc.synth_rhf(o, v, n, p, 80)
    vv = np.zeros((p, p)); vev = np.zeros((p, p))
    lib.check(lib._libagf2.agf2_b200_ctx_build_part_loop_rhf(h, ixq, qja, ei, ea, p, o, v, n, 0, o, 1.0, 1.0, vv, vev))
    rv, rev = orc.np_build_part_loop_rhf(ixq, qja, ei, ea, p, o, v, n, 0, o)
    assert relerr(vv, rv) < TOL and relerr(vev, rev) < TOL
    lib.check(lib._libagf2.agf2_b200_ctx_destroy(h))
    _check_rhf(lib, o, v, n, p, seed=80)                    # default context still alive
    assert lib.comm_info() == (0, 1)
    # sticky status of the void drop-ins: a bad call leaves vv/vev untouched and records the failure
    assert lib.sticky_status(clear=True)[0] == 0
    vv = np.full((p, p), 7.0); vev = np.full((p, p), 7.0)
    lib._libagf2.build_part_loop_rhf(ixq, qja, ei, ea, 0, o, v, n, 0, o, vv, vev)        # nphys = 0
    st, msg = lib.sticky_status(clear=True)
    assert st != 0 and "build_part_loop_rhf" in msg and (vv == 7.0).all() and (vev == 7.0).all()
    assert lib.sticky_status()[0] == 0


@pytest.mark.parametrize("shape", [(12, 30, 100, 42, 1.0, 1.0), (23, 17, 150, 141, 1.2, 1.0 / 3.0), (9, 40, 64, 264, 0.5, 0.0)])
def test_pole_sharded_entry_point_single_rank(lib, shape):
    """The pole-sharded entry point on a context without a communicator (one rank holds the only block): the
    own-block schedule in sub-blocks of one pole, the Coulomb stage on the resident block, `+=` semantics.  The
    multi-rank exchange itself is tests/test_multi_gpu.py (>= 2 GPUs) and tests/test_planner.py (host replay)."""
    import torch
    o, v, n, p, os_f, ss_f = shape
    ixq, qja, ei, ea = orc.synth_rhf(o, v, n, p, 31)
    dev = torch.device("cuda", 0)
    qpad = np.zeros((n, o, v + (v & 1)))
    qpad[:, :, :v] = qja.reshape(n, o, v)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    d_ixq, d_qja, d_ei, d_ea = T(ixq.reshape(o, p, n)), T(qpad), T(ei), T(ea)
    assert lib.pole_block(o, 0, 1) == (0, o)
    rv, rev = orc.np_build_part_loop_rhf(ixq, qja, ei, ea, p, o, v, n, 0, o, os_f, ss_f)
    try:
        for vis in (2 << 30, p * n * 8):
            lib.set_option("visiting_bytes", vis)
            out = torch.full((2, p, p), 0.125, dtype=torch.float64, device=dev)
            lib.moments_rhf_dev_sharded(d_ixq, d_qja, d_ei, d_ea, os_factor=os_f, ss_factor=ss_f, vv=out[0], vev=out[1],
                                        sync=True, ei_host=ei)
            got = out.cpu().numpy() - 0.125
            assert relerr(got[0], rv) < TOL and relerr(got[1], rev) < TOL, (shape, vis, relerr(got[0], rv), relerr(got[1], rev))
            assert lib.last_p2p_bytes() == 0
    finally:
        lib.set_option("visiting_bytes", 2 << 30)
