"""Worker of tests/test_multi_gpu.py: one process per GPU (launched by torch.distributed.run, which only provides
RANK / WORLD_SIZE / LOCAL_RANK / MASTER_PORT -- the collective is the library's own NCCL communicator).
Every rank checks, against the oracle on the same seeded inputs:
  * lib.moments_rhf_multi / moments_uhf_multi  (host pointers: 1/N upload + pipelined all-gather + all-reduce)
  * lib.moments_rhf_dev(comm=True)             (device-resident: part = rank, all-reduce inside the library)
  * the OptRAGF2 driver under N ranks vs the same driver on the oracle backend
and that the results are identical bits on every rank."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

TOL = 1e-10


def relerr(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def main():
    import torch
    from auxgf_b200.lib import agf2 as lib
    from auxgf_b200.util import mpi
    from oracle import agf2_oracle as orc
    comm = mpi.init()
    rank, world = comm.rank, comm.size
    assert world == int(os.environ["WORLD_SIZE"]) and lib.comm_info() == (rank, world)
    dev = torch.device("cuda", torch.cuda.current_device())

    def same_on_all_ranks(a, what):
        t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        hi = lib.comm_allreduce(t.clone(), "max"); lo = lib.comm_allreduce(t.clone(), "min")
        torch.cuda.synchronize()
        assert torch.equal(hi, lo), "%s differs between ranks" % what

    # --- host-pointer path, shapes chosen so that the all-gather route is taken (>= 1 MiB per rank and
    # nocc >= world) and so that it is not (tiny)
    for (o, v, n, p, seed) in ((24, 120, 600, 264, 1), (5, 8, 40, 13, 2), (33, 64, 257, 141, 3)):
        ixq, qja, ei, ea = orc.synth_rhf(o, v, n, p, seed)
        vv, vev = lib.moments_rhf_multi(ixq, qja, ei, ea, p)
        rv, rev = orc.np_build_part_loop_rhf(ixq, qja, ei, ea, p, o, v, n, 0, o)
        assert relerr(vv, rv) < TOL and relerr(vev, rev) < TOL, ("rhf_multi", o, v, n, p, relerr(vv, rv), relerr(vev, rev))
        same_on_all_ranks(np.stack([vv, vev]), "moments_rhf_multi")
        gathered, reduced = lib.last_comm_bytes()
        assert reduced > 0
        vv2, vev2 = lib.moments_rhf_multi(ixq, qja, ei, ea, p, os_factor=1.2, ss_factor=1.0 / 3.0)
        rv, rev = orc.np_build_part_loop_rhf(ixq, qja, ei, ea, p, o, v, n, 0, o, 1.2, 1.0 / 3.0)
        assert relerr(vv2, rv) < TOL and relerr(vev2, rev) < TOL

    args = orc.synth_uhf(20, 40, 19, 41, 300, 150, 4)
    vv, vev = lib.moments_uhf_multi(*args, 150)
    rv, rev = orc.np_build_part_loop_uhf(*args, 150, 20, 19, 40, 41, 300, 0, 20)
    assert relerr(vv, rv) < TOL and relerr(vev, rev) < TOL, ("uhf_multi", relerr(vv, rv), relerr(vev, rev))

    # --- device-resident path with the all-reduce inside the library, twice (bit-identical)
    o, v, n, p = 14, 50, 120, 70
    ixq, qja, ei, ea = orc.synth_rhf(o, v, n, p, 12)
    T = lambda a, sh: torch.from_numpy(np.ascontiguousarray(a)).reshape(sh).to(dev)  # noqa: E731
    d = (T(ixq, (o, p, n)), T(qja, (n, o, v)), T(ei, (o,)), T(ea, (v,)))
    outs = []
    for _ in range(2):
        out = torch.full((2, p, p), 0.5, dtype=torch.float64, device=dev)       # += semantics survive the all-reduce
        lib.moments_rhf_dev(*d, vv=out[0], vev=out[1], sync=True, comm=True, ei_host=ei)
        outs.append(out.cpu().numpy() - 0.5)
    r = orc.np_build_part_loop_rhf(ixq, qja, ei, ea, p, o, v, n, 0, o)
    assert relerr(outs[0][0], r[0]) < TOL and relerr(outs[0][1], r[1]) < TOL
    assert np.array_equal(outs[0], outs[1]), "multi-rank build is not bit-reproducible"
    same_on_all_ranks(outs[0], "moments_rhf_dev(comm=True)")

    # --- pole-sharded build: every rank holds only its block of (ix|Q); the other blocks visit over NVLink in
    # sub-blocks (tiny sub-blocks here so that the double-buffered exchange is exercised), SCS factors, ragged sizes,
    # bit-reproducible and identical on every rank
    for (o, v, n, p, seed, os_f, ss_f, vis) in ((14, 50, 120, 70, 21, 1.0, 1.0, 1 << 30), (37, 21, 130, 141, 22, 1.2, 1.0 / 3.0, 3 * 141 * 130 * 8),
                                                (50, 12, 96, 264, 23, 1.0, 1.0, 264 * 96 * 8), (9, 30, 64, 40, 24, 0.7, 0.0, 1 << 30)):
        if o < world:
            continue
        ixq, qja, ei, ea = orc.synth_rhf(o, v, n, p, seed)
        lo, hi = lib.pole_block(o, rank, world)
        blk = T(ixq.reshape(o, p, n)[lo:hi], (hi - lo, p, n))
        qpad = np.zeros((n, o, v + (v & 1)))                 # device arrays need an even pitch (16-byte TMA strides)
        qpad[:, :, :v] = qja.reshape(n, o, v)
        dq = (T(qpad, qpad.shape), T(ei, (o,)), T(ea, (v,)))
        lib.set_option("visiting_bytes", vis)
        outs = []
        for _ in range(2):
            out = torch.full((2, p, p), 0.25, dtype=torch.float64, device=dev)
            lib.moments_rhf_dev_sharded(blk, *dq, os_factor=os_f, ss_factor=ss_f, vv=out[0], vev=out[1], sync=True, ei_host=ei)
            outs.append(out.cpu().numpy() - 0.25)
        r = orc.np_build_part_loop_rhf(ixq, qja, ei, ea, p, o, v, n, 0, o, os_f, ss_f)
        assert relerr(outs[0][0], r[0]) < TOL and relerr(outs[0][1], r[1]) < TOL, \
            ("sharded", o, v, n, p, relerr(outs[0][0], r[0]), relerr(outs[0][1], r[1]))
        assert np.array_equal(outs[0], outs[1]), "pole-sharded build is not bit-reproducible"
        same_on_all_ranks(outs[0], "moments_rhf_dev_sharded")
        if world > 1 and ss_f != 0.0:
            assert lib.last_p2p_bytes() == (world - 1) * (hi - lo) * p * n * 8, lib.last_p2p_bytes()
    lib.set_option("visiting_bytes", 1 << 30)

    # --- the driver: N ranks on the CUDA backend vs the oracle backend in-process
    from auxgf_b200 import backend
    from auxgf_b200.agf2 import OptRAGF2
    from auxgf_b200.hf import ModelRHF
    from oracle_backend import OracleBackend
    hf = ModelRHF(nao=24, nocc=5, naux=70, seed=1).run()
    g = OptRAGF2(hf, verbose=False)
    g.options["maxiter"] = 2
    g.run()
    prev_b = backend.use(OracleBackend())
    prev_c = mpi.use(mpi._Serial())
    try:
        ref = OptRAGF2(hf, verbose=False)
        ref.options["maxiter"] = 2
        ref.run()
    finally:
        backend.use(prev_b)
        mpi.use(prev_c)
    assert abs(g.e_mp2 - ref.e_mp2) < 1e-8 and abs(g.e_tot - ref.e_tot) < 1e-8, (g.e_tot, ref.e_tot)

    mpi.barrier()
    print("RANK %d OK" % rank, flush=True)
    mpi.finalize()


if __name__ == "__main__":
    try:
        main()
    except BaseException as exc:      # the launcher's banner buries the traceback: say what failed on stdout too
        import traceback
        print("RANK %s FAILED: %r\n%s" % (os.environ.get("RANK", "?"), exc, traceback.format_exc()), flush=True)
        raise
