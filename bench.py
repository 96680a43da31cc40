#!/usr/bin/env python
"""bench.py -- DF-AGF2(None,0) self-energy moment build on B200 (BASELINE.json metric).

A "step" is one full moment build of the synthetic shape (default S: nocc=100, nvir=1000, naux=4000,
nphys=nocc+nvir=1100): the occupied-sector call  build_part_loop(o=nocc, v=nvir)  plus the
virtual-sector call  build_part_loop(o=nvir, v=nocc)  (= one ``OptRAGF2.build()``,
auxgf/agf2/optragf2.py:283-293), F_alg = 1.50e15 flop.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--shape nocc,nvir,naux,nphys] [--no-e2e] [--no-cpu]

  value    : seconds per moment build with the DF tensors already resident in HBM (device C-ABI)
  e2e      : the same build through the reference-facing host-pointer C-ABI
             (agf2_b200_build_part_loop_rhf: host ixq/qja in, host vv/vev out, copies in the timed region)
  roofline : dominant kernel k1_form (integral formation): algorithmic FP64 flop of the (xi|ja) blocks the
             launches form (4 P nvir naux per pole pair) / summed launch time measured with CUDA events,
             against the cuBLAS DGEMM peak measured on this pool's B200 (profiles/fp64_peak_r01.json;
             MEASURED_PEAKS.json carries no FP64 figure)
  cpu_baseline / --impl reference : the reference's own C loop (oracle/_ref/agf2.so, built from
             /root/reference/auxgf/lib/libagf2) on the host cores, on a bounded slice of poles,
             scaled to one build.
N > 1 (torchrun): pair-work items are dealt round-robin to the ranks (part/nparts of the C-ABI), one
NCCL all-reduce of [vv|vev] per sector; total work is fixed -> "scaling": "strong".
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


def f_alg(o, v, n, p):
    return 2.0 * p * o * o * v * n + 4.0 * p * p * o * o * v


def f_k1(o, v, n, p):
    return 2.0 * p * o * o * v * n


def fp64_peak():
    try:
        with open(os.path.join(ROOT, "profiles", "fp64_peak_r01.json")) as f:
            return float(json.load(f)["dgemm_tflops_sustained"]), "measured cublasDgemm 8192^3 (profiles/fp64_peak_r01.json)"
    except Exception:
        return DGEMM_PEAK_FALLBACK, "fallback constant (profiles/fp64_peak_r01.json unreadable)"


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
def cpu_reference_sample(o, v, n, p, budget_s=20.0, keep=None):
    """Times the reference loop on a slice of occupied poles of the occupied-sector call and scales by
    executed work to one full build (work per pole is identical in both sectors: 4PovN + 8P^2ov)."""
    from oracle import agf2_oracle as orc
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(0)
    # scratch of the reference is 3*P*o*v doubles per OpenMP thread (optragf2.c:105-107)
    per_thread_gb = 3.0 * p * o * v * 8 / 1e9
    ixq = (rng.standard_normal((o * p, n), dtype=np.float32) / np.sqrt(n)).astype(np.float64)
    qja = (rng.standard_normal((n, o * v), dtype=np.float32) / np.sqrt(n)).astype(np.float64)
    ei = np.sort(rng.uniform(-2.0, -0.5, o)); ea = np.sort(rng.uniform(0.5, 3.0, v))
    if orc.have_ref():
        kind = "reference"
        orc.load_ref()
        try:
            avail_gb = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE") / 1e9
        except Exception:
            avail_gb = 32.0
        omp = int(max(1, min(cores, (0.5 * avail_gb) // max(per_thread_gb, 1e-9), o)))
        blas = max(1, cores // omp)
        orc.ref_set_threads(omp=omp, blas=blas)
        fn = lambda i0, i1: orc.ref_build_part_loop_rhf(ixq, qja, ei, ea, p, o, v, n, i0, i1)  # noqa: E731
        threads = omp * blas
    else:
        kind = "port"
        omp, blas, threads = 1, cores, cores
        fn = lambda i0, i1: orc.np_build_part_loop_rhf(ixq, qja, ei, ea, p, o, v, n, i0, i1)  # noqa: E731
    t0 = time.perf_counter(); fn(0, omp); t1 = time.perf_counter() - t0        # warm + calibrate
    ni = int(max(omp, min(o, omp * max(1, int(budget_s / max(t1, 1e-3))))))
    t0 = time.perf_counter(); res = fn(0, ni); dt = time.perf_counter() - t0
    if keep is not None:     # inputs + result of the timed slice, for the full-size parity check of the GPU path
        keep.update(ixq=ixq, qja=qja, ei=ei, ea=ea, ni=ni, vv=res[0], vev=res[1])
    per_pole = dt / ni
    build_s = per_pole * (o + v)          # o poles in the occupied sector + v poles in the virtual one
    return {"value": build_s, "unit": "s", "cores": threads, "kind": kind,
            "sample": "poles [0,%d) of the occupied-sector call (o=%d,v=%d,naux=%d,nphys=%d) in %.1f s, "
                      "OMP=%d x BLAS=%d threads, scaled by (nocc+nvir)/%d poles (equal work per pole in both sectors)"
                      % (ni, o, v, n, p, dt, omp, blas, ni),
            "tflops_executed": 2.0 * f_alg(o, v, n, p) / o * ni / dt * 1e-12,
            "tflops_algorithmic": (f_alg(o, v, n, p) + f_alg(v, o, n, p)) / build_s * 1e-12}


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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--shape", default="100,1000,4000,1100")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    args = ap.parse_args()
    ensure_built()
    o, v, n, p = [int(x) for x in args.shape.split(",")]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = "synthetic DF shape nocc=%d nvir=%d naux=%d nphys=%d: occupied + virtual sector moment build" % (o, v, n, p)
    flops_build = f_alg(o, v, n, p) + f_alg(v, o, n, p)

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        vals = []
        cb = None
        for _ in range(max(1, args.warmup > 0) + args.steps):
            cb = cpu_reference_sample(o, v, n, p, budget_s=args.cpu_budget)
            vals.append(cb["value"])
        val = float(np.mean(vals[-args.steps:]))
        cb["value"] = val
        print(json.dumps({
            "impl": "reference", "metric": "agf2_moment_build_seconds_per_iter", "value": val, "unit": "s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": val * 1e3,
            "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload, "l2": "n/a (host cores)",
                       "parallelism": "reference OpenMP loop over poles on %d host threads (no GPU)" % cb["cores"]},
            "tflops": flops_build / val * 1e-12, "cpu_baseline": cb,
            "e2e": {"value": val, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    # ------------------------------------------------------------------ B200 arm
    import torch
    import torch.distributed as dist
    from auxgf_b200.lib import agf2 as lib

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def make(sh, scale, seed):
        # the last dimension is padded to an even pitch (16-byte TMA strides); the logical extent is sh[-1]
        g = torch.Generator(device=dev); g.manual_seed(seed)
        sh = tuple(sh[:-1]) + (sh[-1] + (sh[-1] & 1),)
        return torch.randn(sh, generator=g, device=dev, dtype=torch.float64) * scale

    sc = 1.0 / np.sqrt(n)
    # occupied-sector call: o "occupied" poles, v "virtual"; virtual-sector call: roles swapped
    ixq_o = make((o, p, n), sc, 1); qja_o = make((n, o, v), sc, 2)
    ixq_v = make((v, p, n), sc, 3); qja_v = make((n, v, o), sc, 4)
    e_o = torch.sort(torch.rand(o, device=dev, dtype=torch.float64) * 1.5 - 2.0)[0]
    e_v = torch.sort(torch.rand(v, device=dev, dtype=torch.float64) * 2.5 + 0.5)[0]
    out = torch.zeros((2, 2, p, p), device=dev, dtype=torch.float64)   # [sector][vv|vev]

    def step():
        out.zero_()
        lib.moments_rhf_dev(ixq_o, qja_o, e_o, e_v, part=rank, nparts=world, vv=out[0, 0], vev=out[0, 1], sync=False, naux=n)
        lib.moments_rhf_dev(ixq_v, qja_v, e_v, e_o, part=rank, nparts=world, vv=out[1, 0], vev=out[1, 1], sync=False, naux=n)
        if world > 1:
            dist.all_reduce(out)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def staged_step():
        """One step with the two-stream chunk overlap off and per-launch CUDA events inside the library, so that
        every launch is timed alone on the GPU (with the overlap on, two kernels share the SMs and their elapsed
        times add up to more than the wall time).  Same work as step(); used as the last warm-up step."""
        out.zero_()
        lib.set_overlap(False)
        lib.moments_rhf_dev(ixq_o, qja_o, e_o, e_v, part=rank, nparts=world, vv=out[0, 0], vev=out[0, 1], sync=True, naux=n)
        s1, p1 = lib.last_stage_ms(), lib.last_plan()
        lib.moments_rhf_dev(ixq_v, qja_v, e_v, e_o, part=rank, nparts=world, vv=out[1, 0], vev=out[1, 1], sync=True, naux=n)
        s2, p2 = lib.last_stage_ms(), lib.last_plan()
        lib.set_overlap(True)
        if world > 1:
            dist.all_reduce(out)
        return s1, p1, s2, p2

    staged = None
    for w in range(args.warmup):
        if w == args.warmup - 1:
            staged = staged_step()
        else:
            step()
    sync_all()
    lib.launch_count(reset=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k1_ms = k2_ms = 0.0
    sync_all()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    sync_all()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    launches = lib.launch_count()
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    checksum = float(out.sum().item())
    out_host = out.cpu().numpy().copy()

    # per-stage device time: the last warm-up step (or, with --warmup 0, one more untimed step)
    if staged is None:
        staged = staged_step()
        sync_all()
    st1, pl1, st2, pl2 = staged
    k1_ms, k2_ms, m_ms, c_ms = [a + b for a, b in zip(st1, st2)]
    # (xi|ja) blocks formed by k1_form on this rank: two per PAIR / ASYM item, one per DIAG / OS item
    k1_blocks = [2 * (pl["pair"] + pl["asym"]) + pl["diag"] + pl["os"] for pl in (pl1, pl2)]
    k1_flops_rank = 2.0 * p * n * (k1_blocks[0] * v + k1_blocks[1] * o)
    k1_launches = pl1["k1_launches"] + pl2["k1_launches"]

    # ------------------------------------------------------------------ e2e (host buffers, every rank)
    e2e = None
    if not args.no_e2e:
        host_gb = ((o + v) * p * n + 2 * n * o * v) * 8 / 1e9
        try:
            avail_gb = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE") / 1e9
        except Exception:
            avail_gb = 0.0
        need_gb = host_gb * 1.3 * world + 8          # every rank of the node holds its own host copy
        ok = torch.tensor([1.0 if avail_gb > need_gb else 0.0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() > 0:
            h = [np.ascontiguousarray(ixq_o[:, :, :n].cpu().numpy()).reshape(o * p, n),
                 np.ascontiguousarray(qja_o[:, :, :v].cpu().numpy()).reshape(n, o * v),
                 np.ascontiguousarray(ixq_v[:, :, :n].cpu().numpy()).reshape(v * p, n),
                 np.ascontiguousarray(qja_v[:, :, :o].cpu().numpy()).reshape(n, v * o)]
            eo_h, ev_h = e_o.cpu().numpy(), e_v.cpu().numpy()
            del ixq_o, qja_o, ixq_v, qja_v
            torch.cuda.empty_cache()

            def host_step():
                # the call a user of the reference makes: host DF blocks in, host vv/vev out (per rank: its share
                # of the pair-work items), then the one all-reduce of [vv|vev] over the ranks
                # C host-pointer ABI (ixq streams in pole by pole under the pair loop) on every rank, or -- when
                # that upload would outlast a rank's compute (N = 8 at the S shape) -- 1/N of the inputs per rank +
                # NCCL all-gather over NVLink; one all-reduce of [vv|vev] either way
                r0 = lib.moments_rhf_multi(h[0], h[1], eo_h, ev_h, p)
                r1 = lib.moments_rhf_multi(h[2], h[3], ev_h, eo_h, p)
                return np.stack([r0[0], r0[1], r1[0], r1[1]]).reshape(2, 2, p, p)
            # warm-up of the host path: the (10x cheaper) occupied-sector call only; the staging buffers of the
            # virtual-sector call are then allocated inside the timed region
            lib.moments_rhf_multi(h[0], h[1], eo_h, ev_h, p)
            n_e2e = 1 if ms_per_step > 10e3 else max(1, min(args.steps, 2))
            sync_all()
            t0 = time.perf_counter()
            for _ in range(n_e2e):
                res = host_step()
            sync_all()
            dt_t = torch.tensor([(time.perf_counter() - t0) / n_e2e], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(dt_t, op=dist.ReduceOp.MAX)
            dt = float(dt_t.item())
            chk2 = float(res.sum())
            dev_vs_host = max(float(np.abs(res[k, m] - out_host[k, m]).max() / np.abs(out_host[k, m]).max())
                              for k in range(2) for m in range(2))
            e2e = {"value": dt, "unit": "s", "h2d_bytes_per_step": int(host_gb * 1e9) + 8 * 2 * (o + v),
                   "h2d_note": "upper bound per rank: ranks that take the all-gather route upload 1/N of it",
                   "d2h_bytes_per_step": 4 * p * p * 8, "bytes_are": "per rank", "tflops": flops_build / dt * 1e-12,
                   "checksum_rel_diff_vs_device_path": abs(chk2 - checksum) / max(abs(checksum), 1e-300),
                   "max_rel_diff_vs_device_path": dev_vs_host,
                   "api": "auxgf_b200.lib.agf2.moments_rhf_multi: agf2_b200_build_part_loop_rhf_sharded (host pointers, "
                          "pageable NumPy buffers, ixq upload pipelined under the pair loop) or, when the upload would "
                          "outlast a rank's compute, 1/N upload + NCCL all-gather + agf2_b200_moments_rhf_dev; one "
                          "all-reduce of [vv|vev]"}
        else:
            e2e = {"value": None, "unit": "s", "skipped": "host RAM %.0f GB < %.0f GB needed for the host copies" % (avail_gb, need_gb)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = fp64_peak()
    traffic = traffic_note = None
    try:   # DRAM bytes of one k1_form launch from the committed `ncu --set full` capture
        import csv
        with open(os.path.join(ROOT, "profiles", "r01c_k1_form_ncu.csv")) as f:
            rows = [r for r in csv.reader(f) if r and r[0] == "0"]
        m = {r[2]: (float(r[3]), r[4]) for r in rows}
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        rd, wr = m["dram__bytes_read.sum"], m["dram__bytes_write.sum"]
        traffic = rd[0] * scale[rd[1]] + wr[0] * scale[wr[1]]
        traffic_note = ("dram read+write bytes of one k1_form launch (grid %d) of tools/prof_case.py 64,100,4000,1100, "
                        "profiles/r01c_k1_form_ncu.csv; the path is tensor-bound, the figure only shows that (xi|ja) "
                        "is not spilled to HBM" % int(m["launch__grid_size"][0]))
    except Exception:
        pass
    k1_tf = k1_flops_rank / (k1_ms * 1e-3) * 1e-12 if k1_ms > 0 else None
    k2_flops = flops_build - (f_k1(o, v, n, p) + f_k1(v, o, n, p))
    line = {
        "metric": "agf2_moment_build_seconds_per_iter", "value": ms_per_step * 1e-3, "unit": "s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "l2": "inputs (%.1f GB) exceed the 126 MB L2" % ((o * p * n + n * o * v) * 2 * 8 / 1e9),
                   "parallelism": "pair items round-robin over %d GPU(s) + 1 NCCL all-reduce" % world},
        "tflops": flops_build / (ms_per_step * 1e-3) * 1e-12,
        "frac_of_fp64_peak": flops_build / (ms_per_step * 1e-3) * 1e-12 / (peak * world),
        "fp64_peak_tflops": peak, "fp64_peak_source": peak_src,
        "roofline": {"bound": "tensor", "kernel": "k1_form (FP64 DMMA.8x8x4)", "achieved": k1_tf, "peak": peak,
                     "unit": "TFLOP/s", "frac": (k1_tf / peak) if k1_tf else None, "traffic": traffic,
                     "traffic_note": traffic_note,
                     "launches_per_step": k1_launches, "blocks_per_step": sum(k1_blocks),
                     "flop_per_block": "2 * nphys * nvir * naux (one (xi|ja) block of one pole pair)",
                     "k1_ms_per_step": k1_ms, "k2_ms_per_step": k2_ms, "m_build_ms_per_step": m_ms,
                     "coulomb_ms_per_step": c_ms, "coulomb_factorised": [bool(pl1["factorised"]), bool(pl2["factorised"])],
                     "note": "stage times (CUDA events around every launch, inside the library) from the last warm-up "
                             "step, run with the chunk overlap off so that each launch is alone on the GPU; the timed "
                             "steps overlap the two chunk chains on two streams, where per-launch times would add up "
                             "to more than the wall time",
                     "moment_stage_achieved": k2_flops / world / ((k2_ms + m_ms + c_ms) * 1e-3) * 1e-12
                     if k2_ms > 0 else None},
        "gpu_launches": int(launches), "clocks": clocks, "checksum": checksum,
    }
    if e2e is not None:
        line["e2e"] = e2e
    if not args.no_cpu and world == 1:
        keep = {}
        cb = cpu_reference_sample(o, v, n, p, budget_s=args.cpu_budget, keep=keep)
        try:    # full-size parity: the same pole slice of the same inputs through the host-pointer C-ABI
            g = lib.moments_rhf(keep["ixq"], keep["qja"], keep["ei"], keep["ea"], p, 0, keep["ni"])
            cb["parity_max_rel_diff"] = max(float(np.abs(g[k] - keep[nm]).max() / np.abs(keep[nm]).max())
                                            for k, nm in enumerate(("vv", "vev")))
            cb["parity_note"] = "GPU vs CPU reference on poles [0,%d) of the occupied-sector call at full size" % keep["ni"]
        except Exception as e:   # noqa: BLE001
            cb["parity_max_rel_diff"] = None
            cb["parity_note"] = "not checked: %s" % e
        line["cpu_baseline"] = cb
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
