"""One device-resident sector call on a synthetic shape (profiling driver for ncu)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from auxgf_b200.lib import agf2 as lib

o, v, n, p = [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "64,100,4000,1100").split(",")]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(0)
sc = 1.0 / np.sqrt(n)
# (last dimensions padded to an even pitch: 16-byte TMA strides)
ixq = torch.randn((o, p, n + (n & 1)), generator=g, device=dev, dtype=torch.float64) * sc
qja = torch.randn((n, o, v + (v & 1)), generator=g, device=dev, dtype=torch.float64) * sc
ei = torch.sort(torch.rand(o, device=dev, dtype=torch.float64) * 1.5 - 2.0)[0]
ea = torch.sort(torch.rand(v, device=dev, dtype=torch.float64) * 2.5 + 0.5)[0]
falg = 2.0 * p * o * o * v * n + 4.0 * p * p * o * o * v
for r in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    vv, vev = lib.moments_rhf_dev(ixq, qja, ei, ea, sync=True, naux=n)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    k1, k2, km, kc = lib.last_stage_ms()
    print("shape %s: %.4f s, %.2f TF/s alg | k1 %.2f ms (%.2f TF/s) k2 %.2f ms (%.2f TF/s alg) | M %.2f ms, coulomb %.2f ms | launches %d | checksum %.6e"
          % ((o, v, n, p), dt, falg / dt * 1e-12, k1, 2.0 * p * o * o * v * n / k1 * 1e-9, k2,
             4.0 * p * p * o * o * v / k2 * 1e-9, km, kc, lib.launch_count(reset=True), float(vv.sum())))
