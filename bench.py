#!/usr/bin/env python
"""bench.py -- forward+Jacobian spectra/s of the correlated-k hot path on N B200s.

One "step" = one nemesisfmg-equivalent evaluation per rank on BASELINE.json config 2
(NWAVE 4000, NG 20, NGAS 6, NLAY 100, NPAR 10, NX 60, table 20x15): k-interp + random overlap
(ansb200_gas_opacity) -> path radiance + layer Jacobian (ansb200_radiance) -> state-vector
projection (ansb200_jacobian_project) [-> NCCL all-gather of the KK rows when N > 1].
Weak scaling: every rank evaluates its own geometry (atmosphere state seeded by rank) against a
full replica of the table and the ranks' [spectrum | Jacobian] blocks are all-gathered.

Prints ONE JSON line (see DESIGN.md "Measurement").  `--impl reference` times the CPU oracle port
of the reference path on the host cores instead (the reference itself is pure Python + numba and
is not present on the GPU box).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "forward+Jacobian spectra/s (nemesisfmg-equivalent evaluations per second)"
CFG = dict(nwave=4000, ng=20, npress=20, ntemp=15, ngas=6, nlay=100, nvmr=8, ndust=0, npro=100, nx=60, seed=7)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nwave", type=int, default=CFG["nwave"], help="override NWAVE (experiments only)")
    ap.add_argument("--nx", type=int, default=CFG["nx"])
    ap.add_argument("--cpu-sample-waves", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stage-times", action="store_true", help="also print per-kernel times to stderr")
    return ap.parse_args()


def workload_name(cfg):
    return ("config2: nemesisfmg forward+Jacobian, nadir thermal emission, NWAVE=%d NG=%d NGAS=%d NLAY=%d "
            "NPAR=%d NX=%d, k-table %dx%d" % (cfg["nwave"], cfg["ng"], cfg["ngas"], cfg["nlay"],
                                             cfg["nvmr"] + 2 + cfg["ndust"], cfg["nx"], cfg["npress"], cfg["ntemp"]))


def make_case(cfg, rank_seed=0, nwave=None):
    from archnemesis_dist_b200 import synthetic
    c = synthetic.make_fm_case(nwave=nwave or cfg["nwave"], ng=cfg["ng"], npress=cfg["npress"], ntemp=cfg["ntemp"],
                               ngas=cfg["ngas"], nlay=cfg["nlay"], nvmr=cfg["nvmr"], ndust=cfg["ndust"],
                               npro=cfg["npro"], nx=cfg["nx"], seed=cfg["seed"])
    if rank_seed:
        # another geometry / atmosphere state on the same table: perturb T, amounts and the viewing angle
        rng = np.random.default_rng(1000 + rank_seed)
        c["temp"] = c["temp"] + rng.uniform(-3.0, 3.0, size=c["temp"].shape)
        c["amount"] = c["amount"] * 10.0 ** rng.uniform(-0.2, 0.2, size=c["amount"].shape)
        c["SCALE"] = np.full_like(c["SCALE"], 1.0 / np.cos(np.deg2rad(5.0 + 5.0 * rank_seed)))
        c["EMTEMP"] = c["temp"][c["LAYINC"][:, 0]].reshape(-1, 1).copy()
    return c


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path, on a bounded sample of wavenumbers
# ------------------------------------------------------------------------------------------------
def cpu_forward_jacobian(c, rows, nthreads):
    """The reference chain calc_kg -> k_overlapg -> layer opacity -> thermal_g -> g-sum -> map2pro ->
    map2xvec (oracle restatement) on the wavenumber rows `rows` of case c."""
    from oracle import oracle as orc
    tab = c["tab"]
    K = np.ascontiguousarray(tab["K"][rows])
    k, dkdT = orc.calc_k(K, tab["PRESS"], tab["TEMP"], c["press"], c["temp"], want_grad=True, nthreads=nthreads)
    tau, dk = orc.k_overlap(tab["DELG"], k, c["amount"], dkdT=dkdT, nthreads=nthreads)
    tl, tp, dtl = orc.assemble_opacity(tau, dk, c["gas_slot"], c["NVMR"], c["NPAR"], c["taucon"][rows],
                                       c["dtaucon"][rows], c["LAYINC"], c["SCALE"])
    S, dS, dT = orc.thermal_paths(c["ISPACE"], tab["WAVE"][rows], tl, dtl, c["NVMR"], c["NLAYIN"], c["EMTEMP"],
                                  c["LAYPRESS"], c["LAYINC"], c["TSURF"], c["EMISSIVITY"][rows], c["xfac"][rows],
                                  nthreads=nthreads)
    spec, dspec, dts = orc.g_integrate(S, dS, dT, tab["DELG"])
    inc = orc.included_params(c["xmap"])
    d2 = orc.map2pro(dspec, len(rows), c["NVMR"], c["NDUST"], c["NPRO"], 1, c["NLAYIN"], c["LAYINC"], c["DTE"], c["DAM"],
                     c["DCO"], INCPAR=inc)
    return spec, orc.map2xvec(d2, c["xmap"])


def time_cpu(c, nsample, steps, warmup):
    from oracle import oracle as orc
    # every host thread this process may use (torchrun exports OMP_NUM_THREADS=1: the explicit num_threads
    # clause of the oracle's OpenMP loops is what counts, not the environment default)
    try:
        nthreads = len(os.sched_getaffinity(0))
    except AttributeError:
        nthreads = os.cpu_count() or 1
    nthreads = max(1, nthreads)
    nw = c["tab"]["NWAVE"]
    rows = np.unique(np.linspace(0, nw - 1, min(nsample, nw)).astype(int))
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        cpu_forward_jacobian(c, rows, nthreads)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    t = statistics.mean(times)
    full = t * nw / len(rows)          # the path is independent per wavenumber: scale to the whole spectrum
    return dict(value=1.0 / full, unit="spectra/s", cores=nthreads, kind="port",
                sample="%d of %d wavenumbers (every layer, g-ordinate, gas and state element), oracle C port with "
                       "OpenMP over wavenumbers, scaled by NWAVE/sample; %.2f s per sample pass" % (len(rows), nw, t)), t


def run_reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nsample = args.cpu_sample_waves
    c = make_case(cfg, nwave=max(nsample, 8))      # the sample is generated directly (same recipe, seed)
    cb, t = time_cpu(c, nsample, args.steps, args.warmup)
    # c holds only the sample rows; scale to the full NWAVE of the workload
    value = 1.0 / (t * cfg["nwave"] / c["tab"]["NWAVE"])
    cb["value"] = value
    cb["sample"] = cb["sample"].replace("of %d wavenumbers" % c["tab"]["NWAVE"], "of %d wavenumbers" % cfg["nwave"])
    line = dict(impl="reference", metric=METRIC, value=value, unit="spectra/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 / value, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f64", data="synthetic", config=dict(workload=workload_name(cfg)), cpu_baseline=cb,
                e2e=dict(value=value, unit="spectra/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append((time.perf_counter(), ln.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, ln in self.rows:
            if t < t0 or t > t1 + 0.2:
                continue
            f = [x.strip() for x in ln.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except Exception:
                continue
            for nm, v in zip(names, f[2:]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            for t, ln in self.rows[-3:]:
                f = [x.strip() for x in ln.split(",")]
                try:
                    sm.append(float(f[0])); mx.append(float(f[1]))
                except Exception:
                    pass
        if not sm:
            return None
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_b200(args, cfg):
    import torch
    import torch.distributed as dist
    from archnemesis_dist_b200 import engine, plan

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"       # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    c = make_case(cfg, rank_seed=rank)
    tab = c["tab"]
    hp = engine.HotPath(tab["K"], tab["PRESS"], tab["TEMP"], tab["DELG"], tab["WAVE"])
    ev = engine.Evaluation(press_atm=c["press"], temp=c["temp"], amount=c["amount"], gas_slot=c["gas_slot"],
                           NVMR=c["NVMR"], NPAR=c["NPAR"], LAYINC=c["LAYINC"], SCALE=c["SCALE"], NLAYIN=c["NLAYIN"],
                           EMTEMP=c["EMTEMP"], LAYPRESS=c["LAYPRESS"], taucia=c["taucon"], dtaucon=c["dtaucon"],
                           mode=engine.THERMAL, ISPACE=c["ISPACE"], TSURF=c["TSURF"], EMISSIVITY=c["EMISSIVITY"],
                           xfac=c["xfac"])
    M = plan.fold_projection(c["xmap"], c["LAYINC"], c["NLAYIN"], c["DTE"], c["DAM"], c["DCO"], c["NVMR"], c["NDUST"])
    NW, NX = cfg["nwave"], cfg["nx"]
    gathered = torch.empty((world, NW, NX + 1), dtype=torch.float64, device="cuda") if world > 1 else None
    block = torch.empty((NW, NX + 1), dtype=torch.float64, device="cuda")

    def step_resident(staged, timers=None):
        spec, dx, _ = hp.run(staged, timers)
        if world > 1:
            block[:, 0] = spec[:, 0]
            block[:, 1:] = dx[:, 0, :]
            dist.all_gather_into_tensor(gathered, block)
        return spec, dx

    host_out = torch.empty((NW, NX + 1), dtype=torch.float64, pin_memory=True)

    def step_e2e():
        spec, dx, _ = hp.forward_jacobian(ev, M)          # public API: host arrays in
        if world > 1:
            block[:, 0] = spec[:, 0]
            block[:, 1:] = dx[:, 0, :]
            dist.all_gather_into_tensor(gathered, block)
        if world > 1:
            host_all.copy_(gathered, non_blocking=True)
        else:
            # one contiguous [NWAVE, 1+NX] device block, one DMA into pinned memory (strided D2H copies go
            # through a bounce buffer and a second launch each)
            host_out.copy_(torch.cat([spec[:, :1], dx[:, 0, :]], dim=1), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    host_all = torch.empty((world, NW, NX + 1), dtype=torch.float64, pin_memory=True) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing ------------------------------------------------------------------
    staged = hp.stage(ev, True, M)
    torch.cuda.synchronize()
    for _ in range(args.warmup):
        step_resident(staged)
    K = args.steps
    kt = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    st = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    launches0 = hp.launches
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    t_wall0 = time.perf_counter()
    st[0].record()
    for i in range(K):
        step_resident(staged, kt[i])
        st[i + 1].record()
    barrier()
    t_wall1 = time.perf_counter()
    ms_total = reduce_max(st[0].elapsed_time(st[K]))
    launches = hp.launches - launches0
    k_ms = statistics.mean(a.elapsed_time(b) for a, b in kt)
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None

    # ---- end to end through the public API with host buffers ---------------------------------------
    for _ in range(max(1, args.warmup)):
        step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        step_e2e()
    e1.record()
    barrier()
    e2e_ms = reduce_max(e0.elapsed_time(e1))
    h2d = int(ev.h2d_bytes)
    d2h = int((world if world > 1 else 1) * NW * (NX + 1) * 8)

    if args.stage_times and rank == 0:
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        from archnemesis_dist_b200 import ops
        evs[0].record(); tau, dk = hp.gas_opacity(staged)
        evs[1].record()
        s = staged
        spec, dspec, dts = ops.radiance(s.mode, tau, dk, s.gas_slot, s.taucia, s.taudust, s.tauray, s.dtaucon, s.layinc,
                                        s.scale, s.nlayin, s.emtemp, s.laypress, hp.wave_d, hp.delg_d, s.emissivity,
                                        s.xfac, None, None, None, None, s.ISPACE, s.TSURF, s.NVMR, s.NPAR, True)
        evs[2].record(); ops.jacobian_project(dspec, s.M); evs[3].record()
        torch.cuda.synchronize()
        sys.stderr.write("stage ms: gas_opacity %.3f radiance %.3f project %.3f\n" % (
            evs[0].elapsed_time(evs[1]), evs[1].elapsed_time(evs[2]), evs[2].elapsed_time(evs[3])))

    if rank == 0:
        ms_step = ms_total / K
        value = world * 1e3 / ms_step
        # roofline of the dominant kernel (fused k-interp + overlap), SURVEY.md 8d B_kio
        U = plan.planes_touched(staged.plan_host, cfg["ntemp"])
        plane = cfg["nwave"] * cfg["ng"] * cfg["ngas"] * 8
        b_kio = U * plane + cfg["nwave"] * cfg["ng"] * cfg["nlay"] * 8 * (1 + cfg["ngas"] + 1)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = b_kio / (k_ms * 1e-3) / 1e9
        roof = dict(kernel="ans_koverlap_kernel (ansb200_gas_opacity: fused k-interp + random overlap, gradients)",
                    bound="hbm", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                    traffic=TRAFFIC if cfg["nwave"] == CFG["nwave"] else None,
                    peak_source="measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)",
                    algorithmic_bytes=b_kio, planes_touched=U, kernel_ms=k_ms, share_of_step=k_ms / ms_step,
                    ncu_utilisation=NCU_UTIL if cfg["nwave"] == CFG["nwave"] else None)
        line = dict(metric=METRIC, value=value, unit="spectra/s", n_gpus=world, steps=K, warmup=args.warmup,
                    ms_per_step=ms_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                    data="synthetic",
                    config=dict(workload=workload_name(cfg), sharding="one geometry per rank, table replicated, "
                                "NCCL all-gather of [spectrum|Jacobian] rows" if world > 1 else "single GPU",
                                l2="per-step working set (%.0f MB of table planes + 512 MB of tau/dk) exceeds the "
                                   "126 MB L2; no explicit flush" % (U * plane / 1e6)),
                    e2e=dict(value=world * 1e3 / (e2e_ms / K), unit="spectra/s", h2d_bytes_per_step=h2d,
                             d2h_bytes_per_step=d2h, ms_per_step=e2e_ms / K),
                    gpu_launches=launches, roofline=roof, clocks=clocks)
        if not args.no_cpu_baseline and world == 1:
            cb, _ = time_cpu_sampled(cfg, args.cpu_sample_waves)
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel at config 2, from the
# `ncu --set full` capture in profiles/r01_ncu_full_config2.txt (207.1 MB read = every touched table plane
# once, 467.3 MB written; algorithmic B_kio is 715.5 MB, the remainder of the output was still in L2).
# The same capture says what the kernel IS bound by (it is not HBM): issue slots 47 % busy with 16 resident
# warps per SM stalled on fixed-latency dependencies, shared-memory data pipe 48 %, FP64 pipe 9 %.
TRAFFIC = 674.4e6
NCU_UTIL = dict(issue_slots_pct=47.3, smem_data_pipe_pct=47.6, fp64_pipe_pct=9.2, warps_active_pct=24.1,
                source="profiles/r01_ncu_full_config2.txt (ncu --set full, not taken during the timed run)")


def time_cpu_sampled(cfg, nsample):
    c = make_case(cfg, nwave=max(nsample, 8))
    cb, t = time_cpu(c, nsample, steps=1, warmup=1)
    value = 1.0 / (t * cfg["nwave"] / c["tab"]["NWAVE"])
    cb["value"] = value
    cb["sample"] = cb["sample"].replace("of %d wavenumbers" % c["tab"]["NWAVE"], "of %d wavenumbers" % cfg["nwave"])
    return cb, t


def main():
    args = parse()
    cfg = dict(CFG)
    cfg["nwave"] = args.nwave
    cfg["nx"] = args.nx
    if args.impl == "reference":
        run_reference_arm(args, cfg)
    else:
        run_b200(args, cfg)


if __name__ == "__main__":
    main()
