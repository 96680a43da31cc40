"""GPU, >= 2 devices: the multi-rank paths on real hardware -- one process per GPU, the library's own NCCL
communicator (agf2_b200_comm_*), sharded upload + all-gather, the all-reduce of [vv|vev] inside the library --
against the oracle.  Replaces the mpi4py runs of the reference (auxgf/mpi/agf2/optragf2.py:370-395).
Skipped on a single-GPU box; tools/gpu_multi.sh runs it under `gpurun --gpus 2` and the log is kept in profiles/."""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_multi_rank_nccl_parity(world):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    port = 29600 + (os.getpid() % 300) + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "helpers", "multi_rank_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    ok = len(set(re.findall(r"RANK (\d+) OK", out.stdout)))     # (the ranks' lines may interleave)
    if out.returncode != 0 or ok != world:
        fails = re.findall(r"RANK \S+ FAILED: .*", out.stdout)
        log = os.path.join(ROOT, "gpurun_out")
        if os.path.isdir(log):
            with open(os.path.join(log, "multi_rank_fail_%d.log" % world), "w") as f:
                f.write(out.stdout + "\n==== stderr ====\n" + out.stderr)
        assert False, (fails[:2], out.stdout[-1500:], out.stderr[-1500:])
    assert out.returncode == 0 and ok == world, (out.stdout[-3000:], out.stderr[-3000:])
