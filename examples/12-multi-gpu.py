# One process per GPU:   torchrun --nproc-per-node 8 examples/12-multi-gpu.py     (no MPI, no torch.distributed)
# The NCCL communicator lives inside libagf2_b200.so; auxgf_b200.util.mpi plays the role of auxgf/util/mpi.py.
import numpy as np
import torch

from auxgf_b200.agf2 import OptRAGF2
from auxgf_b200.hf import ModelRHF
from auxgf_b200.lib import agf2 as lib
from auxgf_b200.util import mpi

comm = mpi.init()                      # RANK / WORLD_SIZE / LOCAL_RANK from the launcher
rhf = ModelRHF(nao=24, nocc=5, naux=70, seed=1).run()
gf2 = OptRAGF2(rhf, verbose=False, maxiter=5)   # every build_part: this rank's pair items + ONE all-reduce of [vv|vev]
gf2.run()
if comm.rank == 0:
    print('%d ranks: E(tot) = %.10f' % (comm.size, gf2.e_tot))

# Kernel level, (ix|Q) larger than one GPU: every rank holds only its block of poles, the others visit over NVLink.
o, v, n, p = 64, 96, 256, 160
dev = torch.device('cuda', torch.cuda.current_device())
g = torch.Generator(device=dev); g.manual_seed(0)            # the same (Q|ja) and energies on every rank
qja = torch.randn((n, o, v), generator=g, device=dev, dtype=torch.float64) / np.sqrt(n)
ei = torch.sort(torch.rand(o, generator=g, device=dev, dtype=torch.float64) * 1.5 - 2.0)[0]
ea = torch.sort(torch.rand(v, generator=g, device=dev, dtype=torch.float64) * 2.5 + 0.5)[0]
lo, hi = lib.pole_block(o, comm.rank, comm.size)
g.manual_seed(100 + comm.rank)
ixq_block = torch.randn((hi - lo, p, n), generator=g, device=dev, dtype=torch.float64) / np.sqrt(n)
vv, vev = lib.moments_rhf_dev_sharded(ixq_block, qja, ei, ea, os_factor=6/5, ss_factor=1/3)
if comm.rank == 0:
    print('pole-sharded build: tr(vv) = %.8f, %d bytes sent point-to-point by rank 0' % (float(vv.trace()), lib.last_p2p_bytes()))
mpi.finalize()
