"""One device-resident UHF channel call on a synthetic shape."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from auxgf_b200.lib import agf2 as lib

oa, va, ob, vb, n, p = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "100,1000,99,1001,4000,1100").split(",")]
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(0)
sc = 1.0 / np.sqrt(n)
R = lambda *sh: torch.randn(sh, generator=g, device=dev, dtype=torch.float64) * sc  # noqa: E731
ixq_a, qja_a, qja_b = R(oa, p, n), R(n, oa, va + (va & 1)), R(n, ob, vb + (vb & 1))
E = lambda k, lo, w: torch.sort(torch.rand(k, device=dev, dtype=torch.float64) * w + lo)[0]  # noqa: E731
ei_a, ei_b, ea_a, ea_b = E(oa, -2.0, 1.5), E(ob, -2.0, 1.5), E(va, 0.5, 2.5), E(vb, 0.5, 2.5)
falg = 2.0 * p * oa * n * (oa * va + ob * vb) + 4.0 * p * p * oa * (oa * va + ob * vb)
for r in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    vv, vev = lib.moments_uhf_dev(ixq_a, qja_a, qja_b, ei_a, ei_b, ea_a, ea_b, sync=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("uhf shape %s: %.4f s, %.2f TF/s alg | launches %d | checksum %.6e"
          % ((oa, va, ob, vb, n, p), dt, falg / dt * 1e-12, lib.launch_count(reset=True), float(vv.sum())))
