"""CPU-only: the reference arm of bench.py (`--impl reference`: the reference's own C loop, oracle/_ref/agf2.so, or the
NumPy port when that binary is absent) prints one JSON line with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1",
                          "--shape", "6,12,30,20", "--steps", "1", "--warmup", "1", "--cpu-budget", "0.2"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "agf2_moment_build_seconds_per_iter" and d["unit"] == "s"
    assert d["higher_is_better"] is False and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and abs(d["ms_per_step"] - 1e3 * d["value"]) < 1e-6 * d["ms_per_step"]
    assert "nocc=6 nvir=12 naux=30 nphys=20" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--shape", "6,12,30,20", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_projection_parity_identity_against_the_oracle():
    """bench.projection_parity (the full-size check of pole-sharded builds, where no host can hold the inputs of the
    reference loop) on CPU tensors: the projected evaluation reproduces u^T vv u, u^T vev u of the ORACLE's moments,
    complete and -- with the sampled schedule of `shard_step_limit` -- of the moments summed over exactly the sampled
    pairs (built here from explicit integrals), and it notices a wrong result."""
    import numpy as np
    import torch
    sys.path.insert(0, ROOT)
    import bench
    from auxgf_b200.lib import agf2 as lib
    from oracle import agf2_oracle as orc
    o, v, n, p, os_f, ss_f = 11, 6, 14, 9, 1.2, 1.0 / 3.0
    ixq, qja, ei, ea = orc.synth_rhf(o, v, n, p, 3)
    vv, vev = orc.np_build_part_loop_rhf(ixq, qja, ei, ea, p, o, v, n, 0, o, os_f, ss_f)
    dev = torch.device("cpu")
    T = lambda a, sh: torch.from_numpy(np.ascontiguousarray(a)).reshape(sh)  # noqa: E731
    t_ixq, t_qja, t_ei, t_ea = T(ixq, (o, p, n)), T(qja, (n, o, v)), T(ei, (o,)), T(ea, (v,))
    res = torch.from_numpy(np.stack([vv, vev]))
    par = bench.projection_parity(torch, lib, dev, 1, 0, 0, t_ixq, (0, o), t_qja, t_ei, t_ea, res, n, p, os_f, ss_f)
    assert par["max_rel_diff"] < 1e-12, par
    bad = res.clone(); bad[0, 2, 3] += 1e-6 * float(res[0].abs().max())
    assert bench.projection_parity(torch, lib, dev, 1, 0, 0, t_ixq, (0, o), t_qja, t_ei, t_ea, bad, n, p, os_f, ss_f)["max_rel_diff"] > 1e-9

    # sampled schedule of a 3-rank build with sub-blocks of 2 poles, first sub-block of every ring step only
    class FakeLib:
        shard_cover = staticmethod(lib.shard_cover)

        @staticmethod
        def shard_sub_block(nocc, nphys, ld, nranks, ctx=None):
            return 2
    world, limit = 3, 1
    cover, _ = lib.shard_cover(o, world, 2, limit)
    x = np.einsum("ixq,qja->ijxa", ixq.reshape(o, p, n), qja.reshape(n, o, v))           # X_ij[x, a]
    e = ei[:, None, None] + ei[None, :, None] - ea[None, None, :]
    svv = os_f * np.einsum("ijxa,ijya->xy", x, x); svev = os_f * np.einsum("ijxa,ija,ijya->xy", x, e, x)
    for i in range(o):
        for j in range(i):
            if cover[i, j]:
                a = x[i, j] - x[j, i]
                svv += ss_f * a @ a.T; svev += ss_f * (a * e[i, j][None, :]) @ a.T
    sres = torch.from_numpy(np.stack([svv, svev]))
    # (world > 1 would all-reduce the projected poles through NCCL; the single-process form holds every pole already)
    par = bench.projection_parity(torch, _OneProcess(FakeLib, world), dev, 1, 1, limit, t_ixq, (0, o), t_qja, t_ei, t_ea, sres,
                                  n, p, os_f, ss_f)
    assert par["max_rel_diff"] < 1e-12, par
    assert 0 < cover.sum() < o * (o - 1) // 2


class _OneProcess:
    """projection_parity asks the library for the schedule of `world` ranks; run it in one process with all poles."""

    def __init__(self, fake, world):
        self._fake, self._world = fake, world

    def shard_sub_block(self, nocc, nphys, ld, nranks, ctx=None):
        return self._fake.shard_sub_block(nocc, nphys, ld, self._world)

    def shard_cover(self, nocc, nranks, sb, limit=0):
        return self._fake.shard_cover(nocc, self._world, sb, limit)
