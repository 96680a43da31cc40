#!/usr/bin/env python
"""bench.py -- DF-AGF2(None,0) self-energy moment build on B200 (BASELINE.json metric).

A "step" is one full moment build of the synthetic shape (default S: nocc=100, nvir=1000, naux=4000,
nphys=nocc+nvir=1100): the occupied-sector call  build_part_loop(o=nocc, v=nvir)  plus the
virtual-sector call  build_part_loop(o=nvir, v=nocc)  (= one ``OptRAGF2.build()``,
auxgf/agf2/optragf2.py:283-293), F_alg = 1.50e15 flop.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--shape nocc,nvir,naux,nphys] [--os-factor f] [--ss-factor f] [--no-e2e] [--no-cpu]
                    [--sharded [--sample-limit m]]

  value    : seconds per moment build with the DF tensors already resident in HBM (device C-ABI)
  e2e      : the same build through the reference-facing host-pointer C-ABI
             (agf2_b200_ctx_build_part_loop_rhf: host ixq/qja in, host vv/vev out, copies in the timed region)
  roofline : dominant kernel k1_form (integral formation): algorithmic FP64 flop of the (xi|ja) blocks the
             launches form (2 P nvir naux per block) / summed launch time measured with CUDA events, against the
             cuBLAS DGEMM rate measured in this very run (fp64_peak_tflops)
  cpu_baseline / --impl reference : the reference's own C loop (oracle/_ref/agf2.so, built from
             /root/reference/auxgf/lib/libagf2) on the host cores, on bounded pole slices, both thread layouts
             (OpenMP over poles x serial BLAS, and one pole at a time x threaded BLAS), scaled to one build.
Warm-up steps run 1/32 of the pair-work items of both sectors (they warm clocks, allocations and NCCL; they are not
builds); the last one runs with the two-stream overlap off and per-launch CUDA events to get the per-stage times.
N > 1 (one process per GPU, launched by torch.distributed.run): the pair-work items are dealt round-robin to the
ranks and ONE NCCL all-reduce of [vv|vev] per sector runs inside the library (agf2_b200_comm_*, no torch.distributed);
total work is fixed -> "scaling": "strong".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DGEMM_PEAK_FALLBACK = 35.9   # TFLOP/s, cublasDgemm 8192^3 on this pool (profiles/fp64_peak_r01.json)
WARM_PARTS = 32              # a warm-up step runs 1/32 of the pair-work items


def f_alg(o, v, n, p):
    return 2.0 * p * o * o * v * n + 4.0 * p * p * o * o * v


def f_k1(o, v, n, p):
    return 2.0 * p * o * o * v * n


def make_config(o, v, n, p, os_f, ss_f):
    """Identical in both arms (the driver compares them)."""
    gb = (o * p * n + n * o * v) * 2 * 8 / 1e9
    return {"workload": "synthetic DF shape nocc=%d nvir=%d naux=%d nphys=%d: occupied + virtual sector moment build"
                        % (o, v, n, p),
            "nocc": o, "nvir": v, "naux": n, "nphys": p, "os_factor": os_f, "ss_factor": ss_f,
            "l2": "inputs (%.1f GB) exceed the 126 MB L2: no flush needed between steps" % gb,
            "parallelism": "pair-work items dealt round-robin over the GPUs, one all-reduce of [vv|vev] per sector "
                           "(reference arm: OpenMP/BLAS threads of the host, no GPU)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------
# CPU reference leg (oracle/_ref = the reference's own C loop; else the NumPy port)
# ---------------------------------------------------------------------------------------------
def synth_host(o, v, n, p, seed):
    rng = np.random.default_rng(seed)
    ixq = (rng.standard_normal((o * p, n), dtype=np.float32) / np.sqrt(n)).astype(np.float64)
    qja = (rng.standard_normal((n, o * v), dtype=np.float32) / np.sqrt(n)).astype(np.float64)
    ei = np.sort(rng.uniform(-2.0, -0.5, o)); ea = np.sort(rng.uniform(0.5, 3.0, v))
    return ixq, qja, ei, ea


class CpuReference:
    """The reference loop on pole slices of one sector call, in a chosen thread layout.  Work per pole is the same
    in both sectors (4 P o v N + 8 P^2 o v with the same o v), so a build costs (nocc + nvir) x the time per pole."""

    def __init__(self, os_f=1.0, ss_f=1.0):
        from oracle import agf2_oracle as orc
        self.orc = orc
        self.cores = os.cpu_count() or 1
        self.have_ref = orc.have_ref() and os_f == 1.0 and ss_f == 1.0     # the reference C hard-codes os = ss = 1
        self.os_f, self.ss_f = os_f, ss_f
        if self.have_ref:
            orc.load_ref()
        self.kind = "reference" if self.have_ref else "port"

    def layouts(self, o, v, p):
        """[(omp, blas)]: one pole at a time x threaded BLAS, and OpenMP over poles x serial BLAS (memory-capped:
        3 P o v doubles of scratch per thread, optragf2.c:105-107) -- the two layouts of SURVEY 8d."""
        if not self.have_ref:
            return [(1, self.cores)]
        per_thread_gb = 3.0 * p * o * v * 8 / 1e9
        try:
            avail_gb = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE") / 1e9
        except Exception:
            avail_gb = 32.0
        omp = int(max(1, min(self.cores, (0.5 * avail_gb) // max(per_thread_gb, 1e-9), o)))
        out = [(1, self.cores)]
        if omp > 1:
            out.append((omp, 1))       # (threaded OpenBLAS inside an OpenMP team of > 1 threads may hang: never mixed)
        return out

    def run(self, inputs, p, i0, i1, layout):
        """-> (seconds, vv, vev) of poles [i0, i1) in thread layout (omp, blas)."""
        ixq, qja, ei, ea = inputs
        o, v, n = ei.size, ea.size, ixq.shape[1]
        omp, blas = layout
        if self.have_ref:
            self.orc.ref_set_threads(omp=omp, blas=blas)
            t0 = time.perf_counter()
            vv, vev = self.orc.ref_build_part_loop_rhf(ixq, qja, ei, ea, p, o, v, n, i0, i1)
        else:
            t0 = time.perf_counter()
            vv, vev = self.orc.np_build_part_loop_rhf(ixq, qja, ei, ea, p, o, v, n, i0, i1, self.os_f, self.ss_f)
        return time.perf_counter() - t0, vv, vev


def ensure_built():
    """Fresh checkout: the shared libraries are git-ignored build artefacts -- compile them once."""
    need = [os.path.join(ROOT, "auxgf_b200", "lib", "libagf2_b200.so"),
            os.path.join(ROOT, "oracle", "_build", "liboracle_agf2.so")]
    if all(os.path.exists(f) for f in need):
        return
    if int(os.environ.get("RANK", "0")) == 0:
        import __graft_entry__
        __graft_entry__.build()
    else:                                   # the other ranks wait for rank 0's build
        t0 = time.time()
        while not all(os.path.exists(f) for f in need) and time.time() - t0 < 900:
            time.sleep(2.0)
        time.sleep(2.0)                     # let the linker finish writing


def reference_arm(args, o, v, n, p):
    """--impl reference: the reference's own CPU implementation of the path on the host cores.  Inputs are generated
    once; every thread layout is calibrated once on its smallest slice (one pole per OpenMP thread); each step then
    times one such slice in the best layout whose slice fits the per-step budget.  value = (nocc + nvir) x the best
    seconds per pole seen (mean over the timed steps for the stepped layout)."""
    t_start = time.perf_counter()
    ref = CpuReference(args.os_factor, args.ss_factor)
    inputs = synth_host(o, v, n, p, 0)
    cal = {}
    for lay in ref.layouts(o, v, p):
        ni = min(o, lay[0])
        if cal and min(c["s_per_pole"] for c in cal.values()) * ni > args.cpu_cal_budget:
            continue        # this layout's smallest slice alone would blow the budget (estimate from the layouts seen)
        dt = ref.run(inputs, p, 0, ni, lay)[0]
        cal[lay] = {"poles": ni, "seconds": dt, "s_per_pole": dt / ni}
    affordable = [l for l in cal if cal[l]["seconds"] <= max(args.cpu_budget, min(c["seconds"] for c in cal.values()))]
    step_lay = min(affordable, key=lambda l: cal[l]["s_per_pole"])
    ni = cal[step_lay]["poles"]
    samples = []
    nwarm = 1 if args.warmup > 0 else 0
    for s in range(nwarm + args.steps):
        i0 = (s * ni) % max(1, o - ni + 1)
        samples.append(ref.run(inputs, p, i0, i0 + ni, step_lay)[0] / ni)
    stepped = float(np.mean(samples[nwarm:]))
    best_lay = min(cal, key=lambda l: cal[l]["s_per_pole"])
    per_pole = min(stepped, cal[best_lay]["s_per_pole"])
    val = per_pole * (o + v)
    flops_build = f_alg(o, v, n, p) + f_alg(v, o, n, p)
    cb = {"value": val, "unit": "s", "cores": step_lay[0] * step_lay[1], "kind": ref.kind,
          "sample": "poles of the occupied-sector call (o=%d,v=%d,naux=%d,nphys=%d): each of the %d timed steps = %d "
                    "pole(s) with OMP=%d x BLAS=%d threads (mean %.3f s/pole); every layout calibrated once on one "
                    "pole per OpenMP thread; value = (nocc+nvir) x best s/pole (equal work per pole in both sectors)"
                    % (o, v, n, p, args.steps, ni, step_lay[0], step_lay[1], stepped),
          "layouts": {"omp%d_blas%d" % l: c for l, c in cal.items()},
          "best_layout": "omp%d_blas%d" % best_lay, "s_per_pole": per_pole,
          "tflops_executed": 2.0 * f_alg(o, v, n, p) / o / per_pole * 1e-12,
          "wall_s": time.perf_counter() - t_start}
    print(json.dumps({
        "impl": "reference", "metric": "agf2_moment_build_seconds_per_iter", "value": val, "unit": "s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": val * 1e3,
        "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": make_config(o, v, n, p, args.os_factor, args.ss_factor),
        "tflops": flops_build / val * 1e-12, "cpu_baseline": cb,
        "e2e": {"value": val, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
    return 0


def projection_parity(torch, lib, dev, world, k, limit, ixq_blk, blk, qja, e_i, e_a, res, n, p, os_f, ss_f):
    """Size-independent check of one pole-sharded sector call against an independent evaluation: for a random vector u,
        u^T vv u = (os+ss) sum_ija z_ija^2 - ss sum_ija z_ija z_jia ,   z_ija = sum_Q (u^T ixq_i)[Q] qja[Q, j, a]
    (and the same with the weights e_i + e_j - e_a for vev) costs 1/nphys of the build and needs no integral block: the
    projected (ix|Q) of every rank's poles is all-reduced, z comes from one library-independent torch.matmul (cuBLAS).
    With a sampled call (limit > 0) the exchange-type sum runs over exactly the pairs the sample formed (replayed on the
    host by agf2_b200_shard_cover)."""
    oo, vv_ = e_i.numel(), e_a.numel()
    g = torch.Generator(device=dev); g.manual_seed(77 + k)
    u = torch.randn(p, generator=g, device=dev, dtype=torch.float64)
    y = torch.zeros((oo, n), device=dev, dtype=torch.float64)
    for i0 in range(0, blk[1] - blk[0], 8):
        i1 = min(blk[1] - blk[0], i0 + 8)
        y[blk[0] + i0: blk[0] + i1] = torch.einsum("x,ixq->iq", u, ixq_blk[i0:i1, :, :n])
    if world > 1:
        lib.comm_allreduce(y)
    ld = qja.shape[2]
    z = (y @ qja.reshape(n, oo * ld)).reshape(oo, oo, ld)[:, :, :vv_]          # z[i, j, a]
    mask = None
    if limit > 0:
        sb = lib.shard_sub_block(oo, p, ixq_blk.shape[2], world)
        cov, _ = lib.shard_cover(oo, world, sb, limit)
        mask = torch.from_numpy(cov).to(dev).to(torch.float64)                   # [i, j] = 1 for the pairs formed (i > j)
    c0 = c1 = x0 = x1 = 0.0
    for i0 in range(0, oo, 32):
        i1 = min(oo, i0 + 32)
        zi = z[i0:i1]                                   # [i, j, a]
        zt = z[:, i0:i1].transpose(0, 1)                # [i, j, a] = z[j, i, a]
        e = e_i[i0:i1, None, None] + e_i[None, :, None] - e_a[None, None, :]
        sq = zi * zi
        c0 += float(sq.sum()); c1 += float((sq * e).sum())
        if mask is None:
            ex = zi * zt
        else:
            d = zi - zt
            ex = d * d * mask[i0:i1, :, None]
        x0 += float(ex.sum()); x1 += float((ex * e).sum())
    if mask is None:
        want = ((os_f + ss_f) * c0 - ss_f * x0, (os_f + ss_f) * c1 - ss_f * x1)
    else:
        want = (os_f * c0 + ss_f * x0, os_f * c1 + ss_f * x1)
    got = (float(u @ res[0] @ u), float(u @ res[1] @ u))
    rel = [abs(got[m] - want[m]) / abs(want[m]) for m in range(2)]
    return {"max_rel_diff": max(rel), "vv": rel[0], "vev": rel[1]}


def sharded_arm(args, o, v, n, p, torch, lib, mpi, dev, rank, world, local_rank, t_start):
    """--sharded: the pole-sharded build (agf2_b200_ctx_moments_rhf_dev_sharded).  The two sector calls are generated,
    warmed up, timed and checked one after the other (at the L shape one sector's blocks fill a GPU: 69 GB of (ix|Q) +
    50 GB of (Q|ja) per rank for the virtual sector); a step = occupied-sector call + virtual-sector call, each timed with
    its inputs resident in HBM."""
    os_f, ss_f = args.os_factor, args.ss_factor
    flops_build = f_alg(o, v, n, p) + f_alg(v, o, n, p)

    def sync_all():
        if world > 1:
            lib.comm_barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            lib.comm_allreduce(t, "max")
        return float(t.item())

    sampler_peak = ClockSampler(local_rank).start() if rank == 0 else None
    peaks = measure_fp64_peak(torch, dev, lib)
    peak_clocks = sampler_peak.stop() if rank == 0 else None
    peak = peaks["cublas"]
    sc = 1.0 / np.sqrt(n)
    ge = torch.Generator(device=dev); ge.manual_seed(5)
    e_o = torch.sort(torch.rand(o, generator=ge, device=dev, dtype=torch.float64) * 1.5 - 2.0)[0]
    e_v = torch.sort(torch.rand(v, generator=ge, device=dev, dtype=torch.float64) * 2.5 + 0.5)[0]
    eo_h, ev_h = e_o.cpu().numpy(), e_v.cpu().numpy()
    kw = dict(os_factor=os_f, ss_factor=ss_f, naux=n)
    t_setup = time.perf_counter() - t_start

    sectors = []
    clocks_all = []
    launches = 0
    executed = [0.0, 0.0, 0.0]
    checksum = 0.0
    for k, (oo, vv_, ei, ea, eh, seed, limit) in enumerate(((o, v, e_o, e_v, eo_h, 1, 0), (v, o, e_v, e_o, ev_h, 3, args.sample_limit))):
        lo, hi = lib.pole_block(oo, rank, world)
        ld = n + (n & 1)
        ixq = torch.empty((hi - lo, p, ld), device=dev, dtype=torch.float64)
        g = torch.Generator(device=dev)
        for i in range(lo, hi):                       # every pole from its own seed: any rank could regenerate any pole
            g.manual_seed(seed * 1000003 + i)
            torch.randn((p, ld), generator=g, device=dev, dtype=torch.float64, out=ixq[i - lo])
        ixq.mul_(sc)
        ldv = vv_ + (vv_ & 1)
        g.manual_seed(seed + 1)                       # (Q|ja): the same on every rank
        qja = torch.empty((n, oo, ldv), device=dev, dtype=torch.float64)
        for q0 in range(0, n, 256):
            torch.randn((min(n, q0 + 256) - q0, oo, ldv), generator=g, device=dev, dtype=torch.float64, out=qja[q0:q0 + 256])
        qja.mul_(sc)
        res = torch.zeros((2, p, p), device=dev, dtype=torch.float64)
        sb = lib.shard_sub_block(oo, p, ld, world)

        def call(sync):
            res.zero_()
            lib.moments_rhf_dev_sharded(ixq, qja, ei, ea, vv=res[0], vev=res[1], sync=sync, ei_host=eh, **kw)

        # warm-up: one visiting sub-block per ring step (the whole exchange pattern, a fraction of the pairs); the last one
        # with the two-stream overlap off and per-launch CUDA events (stage times)
        lib.set_option("shard_step_limit", 1)
        stage = plan = None
        for w in range(max(args.warmup, 1)):
            last = w == max(args.warmup, 1) - 1
            if last:
                lib.set_overlap(False)
            call(last)
            if last:
                stage, plan = lib.last_stage_ms(), lib.last_plan()
                lib.set_overlap(True)
        lib.set_option("shard_step_limit", limit)
        sync_all()
        lib.launch_count(reset=True)
        lib.flop_counters(reset=True)
        sampler = ClockSampler(local_rank).start() if rank == 0 else None
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync_all()
        ev0.record()
        for _ in range(args.steps):
            call(False)
        ev1.record()
        sync_all()
        ms = max_over_ranks(ev0.elapsed_time(ev1)) / args.steps
        if rank == 0:
            clocks_all.append(sampler.stop())
        launches += lib.launch_count()
        executed = [a + b / args.steps for a, b in zip(executed, lib.flop_counters())]
        p2p = lib.last_p2p_bytes()
        gathered, reduced = lib.last_comm_bytes()
        checksum += float(res.sum().item())
        par = projection_parity(torch, lib, dev, world, k, limit, ixq, (lo, hi), qja, ei, ea, res, n, p, os_f, ss_f)
        pairs_full = (oo * (oo - 1) // 2 + world - 1) // world
        pairs_run = pairs_full
        ms_full = ms
        coul_ms = stage[2] + stage[3]
        if limit > 0:
            pairs_run = int(lib.shard_cover(oo, world, sb, limit)[1][:, 0].max())
            ms_full = coul_ms + max(ms - coul_ms, 0.0) * pairs_full / max(pairs_run, 1)
        pole_bytes = p * ld * 8
        sectors.append({"nocc": oo, "nvir": vv_, "ms_measured": ms, "ms": ms_full, "sampled_sub_blocks_per_ring_step": limit,
                        "pairs_run_per_rank": pairs_run, "pairs_full_per_rank": pairs_full, "coulomb_stage_ms": coul_ms,
                        "sub_block_poles": sb, "ixq_bytes_per_rank": (hi - lo) * pole_bytes, "qja_bytes_per_rank": qja.numel() * 8,
                        "nvlink_p2p_bytes_sent_per_rank_measured": p2p,
                        "nvlink_p2p_bytes_sent_per_rank_full_build": (world - 1) * (hi - lo) * pole_bytes,
                        "nvlink_allgather_bytes_sent_per_rank": gathered, "nvlink_allreduce_bytes_sent_per_rank": reduced,
                        "projection_parity": par, "stage": stage, "plan": plan})
        del ixq, qja, res
        torch.cuda.empty_cache()

    ms_per_step = sectors[0]["ms"] + sectors[1]["ms"]
    if rank != 0:
        mpi.finalize()
        return 0
    k1_ms = sum(s["stage"][0] for s in sectors)
    k1_flops = sum(2.0 * p * n * 2 * s["plan"]["pair"] * s["nvir"] for s in sectors)
    k1_tf = k1_flops / (k1_ms * 1e-3) * 1e-12 if k1_ms > 0 else None
    sampled = any(s["sampled_sub_blocks_per_ring_step"] > 0 for s in sectors)
    parity = {"max_rel_diff": max(s["projection_parity"]["max_rel_diff"] for s in sectors),
              "occupied_sector": sectors[0]["projection_parity"]["max_rel_diff"],
              "virtual_sector": sectors[1]["projection_parity"]["max_rel_diff"],
              "note": "projection identity u^T vv u, u^T vev u (random u) of the %d-rank pole-sharded result against an "
                      "independent evaluation from the projected (ix|Q) (one torch.matmul per sector), at full size%s; the "
                      "element-wise parity of this code path against the reference is tests/test_multi_gpu.py"
                      % (world, "; virtual sector: over exactly the sampled pairs" if sampled else "")}
    cfg = make_config(o, v, n, p, os_f, ss_f)
    cfg["parallelism"] = "pole-sharded (ix|Q): pole blocks visit over NVLink (ncclSend/ncclRecv, double-buffered under the " \
                         "pair loop), the pairs of a block pair split by parity, M0/M1 all-gathered, one all-reduce of [vv|vev] per sector"
    for s in sectors:
        s.pop("projection_parity"); s["stage_ms"] = dict(zip(("k1_form", "k2_moment", "m_build", "coulomb"), s.pop("stage")))
    line = {
        "metric": "agf2_moment_build_seconds_per_iter", "value": ms_per_step * 1e-3, "unit": "s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
        "tflops": flops_build / (ms_per_step * 1e-3) * 1e-12,
        "algorithmic_tflops_over_peak": flops_build / (ms_per_step * 1e-3) * 1e-12 / (peak * world),
        "extrapolated": ("value = occupied sector (complete) + virtual sector extrapolated from a sample of its pair-work items "
                         "(first %d visiting sub-block(s) of every ring step): Coulomb stage + pair-loop time x pairs_full / "
                         "pairs_run; see `sectors`" % args.sample_limit) if sampled else None,
        "fp64_peak_tflops": peak, "fp64_peak_source": "cublasDgemm 8192^3 (torch.matmul f64) timed for 1.5 s in this run",
        "fp64_peak_clocks": peak_clocks, "inhouse_gemm_tflops": peaks["inhouse"],
        "roofline": {"bound": "tensor", "kernel": "k1_form (FP64 DMMA.8x8x4)", "achieved": k1_tf, "peak": peak,
                     "unit": "TFLOP/s", "frac": (k1_tf / peak) if k1_tf else None, "traffic": None,
                     "flop_per_block": "2 * nphys * nvir * naux (one (xi|ja) block of one pole pair)",
                     "staged_slice": "one visiting sub-block per ring step of both sectors, overlap off, CUDA events around "
                                     "every launch (last warm-up step of each sector)"},
        "gpu_launches": int(launches), "clocks": clocks_all[-1] if clocks_all else None, "clocks_occupied_sector": clocks_all[0] if clocks_all else None,
        "checksum": checksum, "sectors": sectors, "parity": parity,
        "nvlink_bytes_sent_per_rank_per_build": sum(s["nvlink_p2p_bytes_sent_per_rank_full_build"] + s["nvlink_allgather_bytes_sent_per_rank"]
                                                    + s["nvlink_allreduce_bytes_sent_per_rank"] for s in sectors),
        "e2e": {"value": None, "unit": "s", "skipped": "pole-sharded build: the inputs exist only as per-GPU blocks in HBM (the whole "
                                                       "(ix|Q) fits neither one GPU nor the host-pointer drop-in call)"},
        "wall_s": {"setup": t_setup, "total": time.perf_counter() - t_start},
    }
    print(json.dumps(line))
    mpi.finalize()
    return 0


def measure_fp64_peak(torch, dev, lib, seconds=1.5):
    """cuBLAS DGEMM 8192^3 (torch.matmul, f64) back to back for ~`seconds`, and the library's own GEMM on the same
    operands: the roofline denominator, measured beside the result."""
    nn = 8192
    a = torch.randn((nn, nn), device=dev, dtype=torch.float64)
    b = torch.randn((nn, nn), device=dev, dtype=torch.float64)
    c = torch.empty((nn, nn), device=dev, dtype=torch.float64)
    out = {}
    for name, fn in (("cublas", lambda: torch.matmul(a, b, out=c)), ("inhouse", lambda: lib.gemm_dev(a, b, out=c))):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps, t0 = 0, time.perf_counter()
        e0.record()
        while True:
            for _ in range(4):
                fn()
            reps += 4
            torch.cuda.synchronize()
            if time.perf_counter() - t0 > seconds:
                break
        e1.record()
        torch.cuda.synchronize()
        out[name] = 2.0 * nn ** 3 * reps / (e0.elapsed_time(e1) * 1e-3) * 1e-12
    del a, b, c
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--shape", default="100,1000,4000,1100")
    ap.add_argument("--os-factor", type=float, default=1.0)
    ap.add_argument("--ss-factor", type=float, default=1.0)
    ap.add_argument("--sharded", action="store_true",
                    help="pole-sharded build: every rank holds 1/N of (ix|Q), the other blocks visit over NVLink "
                         "(for shapes whose (ix|Q) does not fit one GPU, e.g. --shape 250,2500,10000,2750)")
    ap.add_argument("--sample-limit", type=int, default=0,
                    help="--sharded only: run only the first m visiting sub-blocks of every ring step of the virtual-sector "
                         "call (a bounded sample of a build that takes minutes) and extrapolate by the pair count")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=6.0, help="seconds per CPU sample (timed steps)")
    ap.add_argument("--cpu-cal-budget", type=float, default=60.0, help="seconds for the one-off calibration of a thread layout")
    args = ap.parse_args()
    t_start = time.perf_counter()
    ensure_built()
    o, v, n, p = [int(x) for x in args.shape.split(",")]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    flops_build = f_alg(o, v, n, p) + f_alg(v, o, n, p)
    os_f, ss_f = args.os_factor, args.ss_factor

    if args.impl == "reference":
        if rank != 0:
            return 0
        return reference_arm(args, o, v, n, p)

    # ------------------------------------------------------------------ B200 arm
    import torch
    from auxgf_b200.lib import agf2 as lib
    from auxgf_b200.util import mpi

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    comm = mpi.init()               # the library's NCCL communicator (one rank per process); serial when world == 1
    assert comm.size == world and comm.rank == rank

    if args.sharded:
        return sharded_arm(args, o, v, n, p, torch, lib, mpi, dev, rank, world, local_rank, t_start)

    def sync_all():
        if world > 1:
            lib.comm_barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            lib.comm_allreduce(t, "max")
        return float(t.item())

    sampler_peak = ClockSampler(local_rank).start() if rank == 0 else None
    peaks = measure_fp64_peak(torch, dev, lib)
    peak_clocks = sampler_peak.stop() if rank == 0 else None
    peak = peaks["cublas"]

    def make(sh, scale, seed):
        # the last dimension is padded to an even pitch (16-byte TMA strides); the logical extent is sh[-1]
        g = torch.Generator(device=dev); g.manual_seed(seed)
        sh = tuple(sh[:-1]) + (sh[-1] + (sh[-1] & 1),)
        return torch.randn(sh, generator=g, device=dev, dtype=torch.float64) * scale

    sc = 1.0 / np.sqrt(n)
    # occupied-sector call: o "occupied" poles, v "virtual"; virtual-sector call: roles swapped
    ixq_o = make((o, p, n), sc, 1); qja_o = make((n, o, v), sc, 2)
    ixq_v = make((v, p, n), sc, 3); qja_v = make((n, v, o), sc, 4)
    ge = torch.Generator(device=dev); ge.manual_seed(5)       # (seeded: every rank must hold the same energies)
    e_o = torch.sort(torch.rand(o, generator=ge, device=dev, dtype=torch.float64) * 1.5 - 2.0)[0]
    e_v = torch.sort(torch.rand(v, generator=ge, device=dev, dtype=torch.float64) * 2.5 + 0.5)[0]
    eo_h, ev_h = e_o.cpu().numpy(), e_v.cpu().numpy()
    out = torch.zeros((2, 2, p, p), device=dev, dtype=torch.float64)   # [sector][vv|vev]
    kw = dict(os_factor=os_f, ss_factor=ss_f, naux=n)

    def step():
        # device-resident build: this rank's share of the pair-work items of both sectors; the all-reduce of the packed
        # [vv|vev] runs inside the library on the same stream (comm=True)
        out.zero_()
        lib.moments_rhf_dev(ixq_o, qja_o, e_o, e_v, vv=out[0, 0], vev=out[0, 1], sync=False, comm=True, ei_host=eo_h, **kw)
        lib.moments_rhf_dev(ixq_v, qja_v, e_v, e_o, vv=out[1, 0], vev=out[1, 1], sync=False, comm=True, ei_host=ev_h, **kw)

    nw = WARM_PARTS * world

    def warm_step(w, staged):
        """1/32 of this rank's pair-work items of both sectors.  staged: two-stream overlap off and per-launch CUDA
        events inside the library, so that every launch is timed alone on the GPU (with the overlap on two kernels
        share the SMs and their elapsed times add up to more than the wall time)."""
        out.zero_()
        part = (w % WARM_PARTS) * world + rank
        if staged:
            lib.set_overlap(False)
        lib.moments_rhf_dev(ixq_o, qja_o, e_o, e_v, part=part, nparts=nw, vv=out[0, 0], vev=out[0, 1], sync=staged, **kw)
        s1, p1 = (lib.last_stage_ms(), lib.last_plan()) if staged else (None, None)
        lib.moments_rhf_dev(ixq_v, qja_v, e_v, e_o, part=part, nparts=nw, vv=out[1, 0], vev=out[1, 1], sync=staged, **kw)
        s2, p2 = (lib.last_stage_ms(), lib.last_plan()) if staged else (None, None)
        if staged:
            lib.set_overlap(True)
        if world > 1:
            lib.comm_allreduce(out)
        return s1, p1, s2, p2

    t_setup = time.perf_counter() - t_start
    staged = None
    for w in range(max(args.warmup, 1)):
        last = w == max(args.warmup, 1) - 1
        r = warm_step(w, last)
        if last:
            staged = r
    sync_all()
    t_warm = time.perf_counter() - t_start - t_setup
    lib.launch_count(reset=True)
    lib.flop_counters(reset=True)
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    sync_all()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    launches = lib.launch_count()
    executed = [x / args.steps for x in lib.flop_counters()]     # this rank's executed tensor flops per step
    ms_per_step = max_over_ranks(ms) / args.steps
    comm_bytes_step = lib.last_comm_bytes()[1] * 2 if world > 1 else 0      # two sectors
    checksum = float(out.sum().item())
    out_host = out.cpu().numpy().copy()

    st1, pl1, st2, pl2 = staged
    k1_ms, k2_ms, m_ms, c_ms = [a + b for a, b in zip(st1, st2)]
    # (xi|ja) blocks formed by k1_form in the staged slice: two per PAIR / ASYM item, one per DIAG / OS item
    k1_blocks = [2 * (pl["pair"] + pl["asym"]) + pl["diag"] + pl["os"] for pl in (pl1, pl2)]
    k1_flops_slice = 2.0 * p * n * (k1_blocks[0] * v + k1_blocks[1] * o)
    k1_launches = pl1["k1_launches"] + pl2["k1_launches"]
    k1_tf = k1_flops_slice / (k1_ms * 1e-3) * 1e-12 if k1_ms > 0 else None

    # ------------------------------------------------------------------ e2e (host buffers, every rank)
    e2e = None
    hosts = None
    need_hosts = (not args.no_e2e) or (not args.no_cpu)
    if need_hosts:
        host_gb = ((o + v) * p * n + 2 * n * o * v) * 8 / 1e9
        try:
            avail_gb = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE") / 1e9
        except Exception:
            avail_gb = 0.0
        need_gb = host_gb * 1.15 * (world if not args.no_e2e else 1) + 8    # every rank holds its own host copy
        ok = 1.0 if avail_gb > need_gb else 0.0
        ok = -max_over_ranks(-ok)
        if ok > 0 and (not args.no_e2e or rank == 0):
            hosts = [np.ascontiguousarray(ixq_o[:, :, :n].cpu().numpy()).reshape(o * p, n),
                     np.ascontiguousarray(qja_o[:, :, :v].cpu().numpy()).reshape(n, o * v),
                     np.ascontiguousarray(ixq_v[:, :, :n].cpu().numpy()).reshape(v * p, n),
                     np.ascontiguousarray(qja_v[:, :, :o].cpu().numpy()).reshape(n, v * o)]
        elif ok <= 0:
            e2e = {"value": None, "unit": "s", "skipped": "host RAM %.0f GB < %.0f GB needed for the host copies" % (avail_gb, need_gb)}
    t_e2e0 = time.perf_counter()
    if not args.no_e2e and hosts is not None:
        h = hosts

        def host_step():
            # the call a user of the reference makes: host DF blocks in, host vv/vev out.  One C-ABI call per sector
            # (agf2_b200_ctx_build_part_loop_rhf): every rank uploads 1/N of ixq and qja (ixq in groups of poles,
            # pipelined under the pair loop), NCCL all-gathers replicate them over NVLink, one all-reduce of [vv|vev]
            r0 = lib.moments_rhf_multi(h[0], h[1], eo_h, ev_h, p, os_f, ss_f)
            g0 = lib.last_comm_bytes()
            r1 = lib.moments_rhf_multi(h[2], h[3], ev_h, eo_h, p, os_f, ss_f)
            g1 = lib.last_comm_bytes()
            return np.stack([r0[0], r0[1], r1[0], r1[1]]).reshape(2, 2, p, p), (g0[0] + g1[0], g0[1] + g1[1])
        # warm-up of the host path: the (10x cheaper) occupied-sector call only; the staging buffers of the
        # virtual-sector call are then allocated inside the timed region
        lib.moments_rhf_multi(h[0], h[1], eo_h, ev_h, p, os_f, ss_f)
        sync_all()
        t0 = time.perf_counter()
        res, nvl = host_step()
        sync_all()
        dt = max_over_ranks(time.perf_counter() - t0)
        dev_vs_host = max(float(np.abs(res[k, m] - out_host[k, m]).max() / np.abs(out_host[k, m]).max())
                          for k in range(2) for m in range(2))
        e2e = {"value": dt, "unit": "s", "steps": 1,
               "h2d_bytes_per_step": int(host_gb * 1e9 / world) + 8 * 2 * (o + v), "d2h_bytes_per_step": 4 * p * p * 8,
               "bytes_are": "per rank (every rank uploads 1/N of ixq and qja; pageable NumPy buffers)",
               "nvlink_bytes_sent_per_rank": {"all_gather": nvl[0], "all_reduce": nvl[1]},
               "tflops": flops_build / dt * 1e-12, "max_rel_diff_vs_device_path": dev_vs_host,
               "api": "auxgf_b200.lib.agf2.moments_rhf_multi -> agf2_b200_ctx_build_part_loop_rhf (host pointers in, host "
                      "vv/vev out; upload, all-gather, pair loop, all-reduce and the result copy all inside the call)"}
    t_e2e = time.perf_counter() - t_e2e0

    # ------------------------------------------------------------------ parity at full size + CPU baseline
    # Rank 0 runs the reference's C loop on the first pole(s) of BOTH sector calls (same inputs, host copies); every rank
    # runs the same pole slices on the GPU (device-resident inputs, sharded + all-reduced as in the timed steps).
    cb = None
    parity = None
    t_cpu0 = time.perf_counter()
    if not args.no_cpu:
        nsl = 1
        gpu_slices = []
        for (ix, qj, ei_, ea_, eh) in ((ixq_o, qja_o, e_o, e_v, eo_h), (ixq_v, qja_v, e_v, e_o, ev_h)):
            g = torch.zeros((2, p, p), device=dev, dtype=torch.float64)
            lib.moments_rhf_dev(ix, qj, ei_, ea_, istart=0, iend=nsl, vv=g[0], vev=g[1], sync=True, comm=True, ei_host=eh, **kw)
            gpu_slices.append(g.cpu().numpy())
        if rank == 0 and hosts is not None:
            ref = CpuReference(os_f, ss_f)
            lay = ref.layouts(o, v, p)[0]                    # one pole at a time x threaded BLAS
            secs, diffs = [], []
            for k, (inp, pp) in enumerate((((hosts[0], hosts[1], eo_h, ev_h), p), ((hosts[2], hosts[3], ev_h, eo_h), p))):
                dt_c, rvv, rvev = ref.run(inp, pp, 0, nsl, lay)
                secs.append(dt_c / nsl)
                diffs.append(max(float(np.abs(gpu_slices[k][0] - rvv).max() / np.abs(rvv).max()),
                                 float(np.abs(gpu_slices[k][1] - rvev).max() / np.abs(rvev).max())))
            parity = {"max_rel_diff": max(diffs), "occupied_sector": diffs[0], "virtual_sector": diffs[1],
                      "note": "GPU (device-resident, %d rank(s), all-reduced) vs the CPU %s on poles [0,%d) of the "
                              "occupied-sector call (o=%d,v=%d) and of the virtual-sector call (o=%d,v=%d) at full size"
                              % (world, ref.kind, nsl, o, v, v, o)}
            if world == 1:
                per_pole = float(np.mean(secs))
                cb = {"value": per_pole * (o + v), "unit": "s", "cores": lay[0] * lay[1], "kind": ref.kind,
                      "sample": "pole 0 of the occupied-sector call in %.2f s and pole 0 of the virtual-sector call in "
                                "%.2f s with OMP=%d x BLAS=%d threads, scaled by (nocc+nvir) poles (equal work per pole); "
                                "`--impl reference` times every thread layout and reports the best"
                                % (secs[0], secs[1], lay[0], lay[1]),
                      "tflops_executed": 2.0 * f_alg(o, v, n, p) / o / per_pole * 1e-12,
                      "parity_max_rel_diff": parity["max_rel_diff"], "parity_note": parity["note"]}
    t_cpu = time.perf_counter() - t_cpu0

    if rank != 0:
        mpi.finalize()
        return 0

    traffic = traffic_note = None
    try:   # DRAM bytes of one k1_form launch from the committed `ncu --set full` capture
        import csv
        import glob
        path = sorted(glob.glob(os.path.join(ROOT, "profiles", "r0*_k1_form_ncu.csv")))[-1]
        with open(path) as f:
            rows = [r for r in csv.reader(f) if r and r[0] == "0"]
        m = {r[2]: (float(r[3]), r[4]) for r in rows}
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        rd, wr = m["dram__bytes_read.sum"], m["dram__bytes_write.sum"]
        traffic = rd[0] * scale[rd[1]] + wr[0] * scale[wr[1]]
        traffic_note = ("dram read+write bytes of one k1_form launch (grid %d CTAs), %s: the DF blocks of the launch's pole "
                        "pairs (ixq rows re-read once per column block, L2 hit rate %.0f %%) plus the workspace write-back; the "
                        "path is tensor-bound (%.0f %% of HBM peak)"
                        % (int(m["launch__grid_size"][0]), os.path.basename(path), m["lts__t_sector_hit_rate.pct"][0],
                           m["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"][0]))
    except Exception:
        pass
    # executed FP64 tensor flops of one build (what the DMMA pipe issues, padding included) from the staged slice
    line = {
        "metric": "agf2_moment_build_seconds_per_iter", "value": ms_per_step * 1e-3, "unit": "s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": make_config(o, v, n, p, os_f, ss_f),
        "tflops": flops_build / (ms_per_step * 1e-3) * 1e-12,
        "algorithmic_tflops_over_peak": flops_build / (ms_per_step * 1e-3) * 1e-12 / (peak * world),
        "algorithmic_note": "F_alg of SURVEY 8d (2Po^2vN + 4P^2o^2v per sector) / time / measured DGEMM rate: NOT a hardware "
                            "fraction -- pair symmetry and the Coulomb factorisation execute fewer flops than F_alg",
        "executed_tflops": sum(executed) * world / (ms_per_step * 1e-3) * 1e-12,
        "executed_frac": sum(executed) / (ms_per_step * 1e-3) * 1e-12 / peak,
        "executed_note": "FP64 tensor flops the launches of rank 0 issued per step (tile padding included; k1_form %.3e, "
                         "k2_moment %.3e, GEMM %.3e) x n_gpus / time; frac = per-GPU rate / measured DGEMM rate"
                         % tuple(executed),
        "fp64_peak_tflops": peak, "fp64_peak_source": "cublasDgemm 8192^3 (torch.matmul f64) timed for 1.5 s in this run",
        "fp64_peak_clocks": peak_clocks, "inhouse_gemm_tflops": peaks["inhouse"],
        "roofline": {"bound": "tensor", "kernel": "k1_form (FP64 DMMA.8x8x4)", "achieved": k1_tf, "peak": peak,
                     "unit": "TFLOP/s", "frac": (k1_tf / peak) if k1_tf else None, "traffic": traffic,
                     "traffic_note": traffic_note,
                     "flop_per_block": "2 * nphys * nvir * naux (one (xi|ja) block of one pole pair)",
                     "staged_slice": {"what": "part %d of %d of the pair-work items of both sectors, overlap off, CUDA events "
                                              "around every launch inside the library (last warm-up step)" % (rank, nw),
                                      "k1_launches": k1_launches, "k1_blocks": sum(k1_blocks), "k1_ms": k1_ms, "k2_ms": k2_ms,
                                      "m_build_ms": m_ms, "coulomb_ms": c_ms,
                                      "coulomb_factorised": [bool(pl1["factorised"]), bool(pl2["factorised"])]},
                     "k1_share_of_staged_slice": k1_ms / max(k1_ms + k2_ms + m_ms + c_ms, 1e-9)},
        "gpu_launches": int(launches), "clocks": clocks, "checksum": checksum,
        "nvlink_bytes_sent_per_rank_per_step": comm_bytes_step,
        "wall_s": {"setup": t_setup, "warmup": t_warm, "timed": ms * 1e-3, "e2e_leg": t_e2e, "cpu_leg": t_cpu,
                   "total": time.perf_counter() - t_start},
    }
    if e2e is not None:
        line["e2e"] = e2e
    if parity is not None:
        line["parity"] = parity
    if cb is not None:
        line["cpu_baseline"] = cb
    print(json.dumps(line))
    mpi.finalize()
    return 0


if __name__ == "__main__":
    sys.exit(main())
