#!/usr/bin/env python
"""bench.py -- trajectories/sec of the batched stiff-ODE hot path (BASELINE.json metric) on N B200s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--sets S] [--impl reference]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Workload (config.workload): BASELINE.json configs[1] -- examples/MAPK.xml parameter scan, 10^6 random kinetic-
parameter sets PER GPU (weak scaling: rank r integrates sets [r*S, (r+1)*S) of one global Philox-keyed batch),
rtol 1e-6 / atol 1e-9, t = 0..1000, 100 output points, the 8 species stored.  One "step" = one pass of the hot
path over that batch (one kernel launch per GPU).

  value      whole-job trajectories/s with inputs resident in HBM (device-generated), timed with CUDA events on the
             launching stream around exactly K launches, barrier + synchronise on both sides, max over ranks
  e2e        the same metric through the C-ABI call with HOST buffers (ODESBatch_runHost: pinned host parameter rows
             in, trajectories out, chunked H2D / kernel / D2H overlap), wall clock, max over ranks
  roofline   dominant kernel odes_bdf_kernel against the FP64 FMA peak measured live on this GPU (MEASURED_PEAKS.json
             has no FP64 figure), plus the HBM view of the algorithmic output bytes
  cpu_baseline  the reference's own CVODE (oracle/_ref, vendored SUNDIALS 2.1.1 compiled unmodified) + the emitted C
             right-hand side, on all host cores, on a bounded sample of the same seeded sets

--impl reference times only that CPU path (rank 0; other ranks exit 0).
"""
import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SEED = 20261019
METRIC = "trajectories/sec (MAPK param scan, rtol 1e-6)"
UNIT = "trajectories/s"


def philox_sets(nominal, first, count, lo=-1.0, hi=1.0):
    """Host restatement of odes_generate_values_kernel (Philox4x32-10, counter = (set, j), key = SEED)."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    nv = len(nominal)
    sets = np.arange(first, first + count, dtype=np.uint64)[:, None] + np.zeros((1, nv), dtype=np.uint64)
    c = [(sets & 0xFFFFFFFF), (sets >> np.uint64(32)), np.arange(nv, dtype=np.uint64)[None, :] + np.zeros((count, 1), dtype=np.uint64),
         np.zeros((count, nv), dtype=np.uint64)]
    k0, k1 = SEED & 0xFFFFFFFF, SEED >> 32
    for _ in range(10):
        p0, p1 = c[0] * np.uint64(M0), c[2] * np.uint64(M1)
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & np.uint64(0xFFFFFFFF), p1 >> np.uint64(32), p1 & np.uint64(0xFFFFFFFF)
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    bits = (c[0] << np.uint64(32)) | c[1]
    u = (bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)
    return nominal[None, :] * np.exp2(lo + (hi - lo) * u)


def count_ops(src):
    """Arithmetic operations (+ - * / and function calls, one each) in the emitted ode_f and jacobi_f bodies."""
    out = {}
    for fn in ("ode_f", "jacobi_f"):
        m = re.search(r"DEV void %s\(.*?\n\{(.*?)\n\}\n" % fn, src, re.S)
        body = re.sub(r"\(void\)\w+;", "", m.group(1))
        stmts = [l.split("=", 1)[1] for l in body.split("\n") if "=" in l and not l.strip().startswith("double")]
        txt = " ".join(stmts)
        txt = re.sub(r"\[[0-9]+\]", "", txt)                         # indices
        txt = re.sub(r"[0-9.]+e[-+][0-9]+", "1", txt)                # exponents of literals
        out[fn] = len(re.findall(r"[-+*/]", txt)) + len(re.findall(r"\b(pow|exp|log|sqrt|sin|cos|tan|odes_\w+)\(", txt))
    return out["ode_f"], out["jacobi_f"]


def algorithmic_flops(neq, Ff, FJ, mean_stats, qbar=4.0):
    """F = nfe*F_f + nje*F_J + nsetups*(2n^3/3 + 2n^2) + nni*(2n^2 + 6n) + nst*n*(q(q+1)/2 + 2(q+1) + 10)."""
    nst, nfe, nsetups, nje, nni = (mean_stats[k] for k in ("nst", "nfe", "nsetups", "nje", "nni"))
    n = neq
    return nfe * Ff + nje * FJ + nsetups * (2 * n ** 3 / 3 + 2 * n * n) + nni * (2 * n * n + 6 * n) + nst * n * (qbar * (qbar + 1) / 2 + 2 * (qbar + 1) + 10)


def bind_to_gpu_numa_node(index):
    """Pins this rank to the CPUs of its GPU's NUMA node before anything is allocated, so that the pinned host
    buffers of the end-to-end leg are node-local (8 ranks x 6.5 GB of results per step otherwise cross the sockets).
    Returns a short description for the JSON line, or None when the topology cannot be read."""
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(index)],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if not bus:
            return None
        bus = bus[-12:]                                              # 00000000:1B:00.0 -> 0000:1b:00.0
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return f"numa node {node}, {len(cpus)} cpus"
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md, clocks line)."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def cpu_model_name():
    try:
        for l in open("/proc/cpuinfo"):
            if l.startswith("model name"):
                return l.split(":", 1)[1].strip()
    except Exception:
        pass
    return None


def status_histogram(st):
    """failed trajectories by CVODE flag"""
    vals, cnt = np.unique(np.asarray(st)[np.asarray(st) != 0], return_counts=True)
    return {str(int(v)): int(c) for v, c in zip(vals, cnt)}


def flops_for(b, m, stats_mean):
    Ff, FJ = count_ops(b.source())
    return algorithmic_flops(m.neq, Ff, FJ, stats_mean), Ff, FJ


def other_config_line(H, so, name, n, fp64_peak, cores, device):
    """One parity-test configuration of BASELINE.json (configs[2..4]) at n sets: device-resident trajectories/s, roofline
    fraction with the oracle's work counters, and the parity of a sample against the reference's vendored CVODE."""
    w = H.workload(name)
    m = w["model"]
    b = so.Batch(m, w["opt"], w["vary"], store=list(range(m.neq)), device=device)
    vals = H.workload_values(w, n, seed=77)
    b.set_values(vals)
    b.integrate(n)
    ms = min(b.integrate(n) for _ in range(2))
    st = b.status(n)
    stats = b.stats(n)
    ns = min(n, 256 if name == "huang" else 512)
    backend = H.REF if H.have_ref() else H.PORT
    cso = H.compile_c_model(m, name)
    t0 = time.perf_counter()
    ref = H.oracle_run(m, w["opt"], w["vary"], vals[:ns], w["nout"], backend=backend, threads=cores, compiled_so=cso)
    dt = time.perf_counter() - t0
    gpu = b.results(ns)
    rep = H.parity_report(gpu, np.transpose(ref["out"][:, :, :m.neq], (1, 2, 0)), stats[:, :ns].T, ref["stats"], w["rtol"], w["atol"])
    fired_same = bool(np.array_equal(b.event_log(ns).T, ref["fired"])) if name == "events" else None
    omean = dict(zip("nst nfe nsetups nje nni ncfn netf".split(), [float(x) for x in ref["stats"].mean(axis=0)]))
    F, Ff, FJ = flops_for(b, m, omean)
    achieved = F * n / (ms * 1e-3) / 1e12
    info = b.kernel_info()
    line = {"workload": name, "baseline_config": {"repressilator": "configs[2]", "events": "configs[3]", "huang": "configs[4]", "goodwin": "configs[4]"}[name],
            "sets": n, "neq": m.neq, "output_points": w["nout"], "t_end": w["end"], "value": n / (ms * 1e-3), "unit": UNIT, "kernel_ms": ms,
            "failed_sets": int((st != 0).sum()), "layout": "warp per trajectory" if "-DODES_WARP=1" in b.source().split("\n")[0] else "trajectory per lane", "kernel_info": info,
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak if fp64_peak else None, "flop_per_trajectory": F},
            "oracle_mean_stats": omean,
            "parity_on_sample": {"sets": ns, "checker": "vendored CVODE 2.3.0 (oracle/_ref)" if backend == H.REF else "restated port", "bit_identical": bool(np.array_equal(gpu, np.transpose(ref["out"][:, :, :m.neq], (1, 2, 0)), equal_nan=True)),
                                 "frac_within_10rtol": rep["frac_within"], "frac_same_step_sequence": rep["frac_same_path"], "max_ratio": rep["max_ratio"], "event_firing_order_identical": fired_same},
            "cpu_traj_per_s": ns / dt, "cpu_cores": cores}
    b.close()
    return line


def scaled_config_line(H, so, name, n, device, world, rank, barrier, max_over_ranks):
    """BASELINE configs[4] "at 2/4/8 GPUs": every rank integrates n sets of the configuration on its GPU (range sharding, no
    collective); trajectories/s of the whole job from the slowest rank's device time.  Parity of this configuration is
    checked at N = 1 (other_configs)."""
    w = H.workload(name)
    m = w["model"]
    b = so.Batch(m, w["opt"], w["vary"], store=list(range(m.neq)), device=device)
    vals = H.workload_values(w, n, seed=77 + rank)
    b.set_values(vals)
    b.integrate(n)
    barrier()
    ms = max_over_ranks(min(b.integrate(n) for _ in range(2)))
    failed = int((b.status(n) != 0).sum())
    info = b.kernel_info()
    warp = "-DODES_WARP=1" in b.source().split("\n")[0]
    b.close()
    return {"workload": name, "baseline_config": "configs[4]", "sets_per_gpu": n, "n_gpus": world, "value": world * n / (ms * 1e-3), "unit": UNIT,
            "kernel_ms_max_over_ranks": ms, "failed_sets_rank0": failed, "layout": "warp per trajectory" if warp else "trajectory per lane", "kernel_info": info}


def entry_leg(H, so, L, w, n, steps):
    """The reference-shaped one-shot entry: IntegratorInstance_integrateBatch with the caller's malloc'd (pageable) arrays and
    ALL model values per row (what the reference's batch loop stores per design point, src/odeSolver.c:313-330)."""
    import ctypes as C
    m = w["model"]
    nv, nrows = len(w["vary"]), w["nout"] + 1
    ii = L.IntegratorInstance_create(m.om, w["opt"])
    vis = (C.c_void_p * nv)(*[L.ODEModel_getVariableIndexByNum(m.om, i) for i in w["vary"]])
    vals = np.ascontiguousarray(philox_sets(w["nominal"], 0, n))
    out = np.zeros((n, nrows, m.nvalues))                       # malloc'd, pageable; touched
    st = np.zeros(n, dtype=np.int32)
    call = lambda: L.IntegratorInstance_integrateBatch(ii, n, nv, vis, vals.ctypes.data, out.ctypes.data, st.ctypes.data)
    t0 = time.perf_counter(); failed0 = call(); first_s = time.perf_counter() - t0       # builds and caches the plan, buffers, ring
    c1 = so.counters()
    t0 = time.perf_counter()
    for _ in range(steps):
        failed = call()
    dt = (time.perf_counter() - t0) / steps
    c2 = so.counters()
    assert failed0 == 0 and failed == 0, (failed0, failed)
    reuse = {k: c2[k] - c1[k] for k in ("nvrtc_compiles", "cubin_cache_loads", "module_loads", "device_allocations", "plans_created", "plan_cache_hits", "kernel_launches")}
    # the same shape through ODESBatch_runHost with pinned buffers, for the comparison
    b = so.Batch(m, w["opt"], w["vary"], device=0)
    rows = b.nrows * b.nstore
    hv, hr, hs = L.ODES_hostAlloc(n * nv * 8), L.ODES_hostAlloc(rows * n * 8), L.ODES_hostAlloc(n * 4)
    pv = np.ctypeslib.as_array(C.cast(hv, C.POINTER(C.c_double)), shape=(n, nv)); pv[:] = vals
    pr = np.ctypeslib.as_array(C.cast(hr, C.POINTER(C.c_double)), shape=(n, b.nrows, b.nstore))
    ps = np.ctypeslib.as_array(C.cast(hs, C.POINTER(C.c_int)), shape=(n,))
    b.run_host(pv, pr, ps)
    t0 = time.perf_counter()
    for _ in range(steps):
        b.run_host(pv, pr, ps)
    dt_pinned = (time.perf_counter() - t0) / steps
    same = bool(np.array_equal(pr, out, equal_nan=True))
    L.ODES_hostFree(hv); L.ODES_hostFree(hr); L.ODES_hostFree(hs)
    b.close()
    L.IntegratorInstance_free(ii)
    d2h = n * nrows * m.nvalues * 8 + n * 4
    return {"entry": "IntegratorInstance_integrateBatch", "value": n / dt, "unit": UNIT, "sets": n, "values_per_row": m.nvalues, "ms_per_call": 1e3 * dt,
            "first_call_s": first_s, "h2d_bytes_per_call": n * nv * 8, "d2h_bytes_per_call": d2h, "d2h_gbs": d2h / dt / 1e9,
            "host_arrays": "malloc'd (pageable), staged through the plan's pinned ring", "second_call_work": reuse,
            "runhost_pinned_same_shape": {"value": n / dt_pinned, "ms_per_call": 1e3 * dt_pinned, "d2h_gbs": d2h / dt_pinned / 1e9},
            "ratio_to_runhost_pinned": dt_pinned / dt, "identical_to_runhost": same,
            "bound": "device-to-host copy of %d values per row: %.1f GB per call" % (m.nvalues, d2h / 1e9)}


def fmad_variant(H, so, L, w, n, fp64_peak, device, cores):
    """NOT the headline: the same kernel compiled with --fmad=true (ODES_FMAD=1).  Contraction into fused multiply-adds
    changes the last bits, so trajectories follow the tolerance gate instead of the reference's exact steps -- the
    difference to the headline value is the price of bit-exactness."""
    m = w["model"]
    os.environ["ODES_FMAD"] = "1"
    try:
        b = so.Batch(m, w["opt"], w["vary"], store=list(range(m.neq)), device=device)
        b.reserve(n)
        nv = len(w["vary"])
        b.generate_values(n, SEED, 0, w["nominal"], w["span"][0] * np.ones(nv), w["span"][1] * np.ones(nv), log10=w["log10"])
        b.integrate(n)
        ms = min(b.integrate(n) for _ in range(2))
        ns = 4096
        vals = b.get_values(ns)
        ref = H.oracle_run(m, w["opt"], w["vary"], vals, w["nout"], backend=H.REF if H.have_ref() else H.PORT, threads=cores, compiled_so=H.compile_c_model(m, "mapk"))
        rep = H.parity_report(b.results(ns), np.transpose(ref["out"][:, :, :m.neq], (1, 2, 0)), b.stats(ns).T, ref["stats"], w["rtol"], w["atol"])
        out = {"value": n / (ms * 1e-3), "unit": UNIT, "kernel_ms": ms, "build": "--fmad=true (ODES_FMAD=1)", "headline": False,
               "parity_on_sample": {"sets": ns, "frac_within_10rtol": rep["frac_within"], "frac_same_step_sequence": rep["frac_same_path"], "max_ratio": rep["max_ratio"]}}
        b.close()
        return out
    finally:
        os.environ.pop("ODES_FMAD", None)


def reference_arm(args, rank):
    """The reference's CPU implementation of the path (vendored CVODE 2.3.0 via oracle/_ref when it was built in the
    container, else the restated port) with all host threads, on bounded samples of the same seeded workload."""
    if rank != 0:
        return
    import helpers as H
    import sbml_odesolver_b200 as so
    w = H.workload("mapk")
    m = w["model"]
    cores = os.cpu_count() or 1
    kind = "reference" if H.have_ref() else "port"
    backend = H.REF if kind == "reference" else H.PORT
    cso = H.compile_c_model(m, "mapk")
    per_step = args.ref_sets or max(2048, min(4000 * cores, 400_000))
    times = []
    for step in range(args.warmup + args.steps):
        vals = philox_sets(w["nominal"], step * per_step, per_step)
        t0 = time.perf_counter()
        r = H.oracle_run(m, w["opt"], w["vary"], vals, w["nout"], backend=backend, threads=cores, compiled_so=cso)
        dt = time.perf_counter() - t0
        assert (r["status"] == 0).all()
        if step >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = per_step * len(times) / total
    sample = f"{per_step} sets/step x {len(times)} steps of the seeded MAPK batch (Philox seed {SEED}), {cores} threads, gcc -O2 -ffp-contract=off, emitted C rhs + analytic Jacobian"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "MAPK.xml parameter scan, rtol 1e-6 atol 1e-9, t=0..1000, 100 output points (CPU: bounded sample per step)",
                       "sets_per_step": per_step, "nvary": len(w["vary"])},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--sets", type=int, default=0, help="trajectories per GPU per step (default 10^6; 10^5 with --sens)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="mapk", choices=["mapk", "repressilator", "goodwin", "huang"],
                    help="mapk = the headline configuration (BASELINE configs[1]); the others are the parity-test configurations, for profiles/")
    ap.add_argument("--ref-sets", type=int, default=0, help="sets per step of the CPU arm (default: scaled to the core count)")
    ap.add_argument("--cpu-sets", type=int, default=0, help="sample size of the cpu_baseline leg (default: scaled to the core count)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-other", action="store_true", help="skip the other_configs block (BASELINE configs[2..4] at 10^5 sets), the e2e_entry leg and the FMA variant")
    ap.add_argument("--entry-sets", type=int, default=200_000, help="sets of the e2e_entry leg (IntegratorInstance_integrateBatch, all model values, malloc'd arrays)")
    ap.add_argument("--sens", default="", help="forward sensitivities (SURVEY 8f rank 4): comma-separated ids, or 'all' for every constant; "
                                               "not the headline -- a line for profiles/ with the vendored CVODES timed beside it")
    args = ap.parse_args()
    if args.sets <= 0:
        args.sets = 100_000 if args.sens else 1_000_000
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return reference_arm(args, rank)

    import helpers as H
    import sbml_odesolver_b200 as so
    L = so.lib()
    if L.ODES_deviceCount() <= local:
        raise SystemExit("bench.py: no CUDA device for this rank -- the batched integrator has no CPU fallback")

    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sens_ids = None if args.sens in ("", "all") else args.sens.split(",")
    w = H.workload(args.workload, sensitivity=1, sens_ids=sens_ids) if args.sens else H.workload(args.workload)
    m, nv, n = w["model"], len(w["vary"]), args.sets
    store = list(range(m.neq))
    b = so.Batch(m, w["opt"], w["vary"], store=store, device=local)
    b.reserve(n)
    lo, hi = w["span"][0] * np.ones(nv), w["span"][1] * np.ones(nv)
    first = rank * n
    b.generate_values(n, SEED, first, w["nominal"], lo, hi, log10=w["log10"])
    fp64_peak = L.ODES_measureFP64Peak(local, 5) if rank == 0 else 0.0

    # ---------------- device-resident timing ----------------
    for _ in range(args.warmup):
        b.integrate(n)
    launches0 = L.ODESBatch_getNumLaunches(b.h)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    L.ODESBatch_synchronize(b.h)
    L.ODESBatch_timerStart(b.h)
    for _ in range(args.steps):
        b.integrate(n, sync=False)
    ms = L.ODESBatch_timerStop(b.h)
    L.ODESBatch_synchronize(b.h)
    barrier()
    clocks = sampler.summary()
    launches = L.ODESBatch_getNumLaunches(b.h) - launches0
    ms_total = max_over_ranks(ms)
    ms_step = ms_total / args.steps
    value = world * n / (ms_step * 1e-3)
    kernel_ms = ms / args.steps                                     # rank-local average launch duration
    st = b.status(n)
    stats = b.stats(n)
    failed = int((st != 0).sum())
    mean_stats = dict(zip("nst nfe nsetups nje nni ncfn netf".split(), [float(x) for x in stats.mean(axis=1)]))

    # ---------------- end to end through the C ABI with host buffers ----------------
    e2e = None
    if not args.no_e2e:
        rows = b.nrows * b.nstore
        hv = L.ODES_hostAlloc(n * nv * 8)
        hr = L.ODES_hostAlloc(rows * n * 8)
        hs = L.ODES_hostAlloc(n * 4)
        if not (hv and hr and hs):
            raise SystemExit("bench.py: pinned host allocation failed")
        import ctypes as C
        vals = np.ctypeslib.as_array(C.cast(hv, C.POINTER(C.c_double)), shape=(n, nv))
        res = np.ctypeslib.as_array(C.cast(hr, C.POINTER(C.c_double)), shape=(n, b.nrows, b.nstore))     # [set][row][value]
        hst = np.ctypeslib.as_array(C.cast(hs, C.POINTER(C.c_int)), shape=(n,))
        hsens, hsp = None, None
        if b.nsens:
            hsp = L.ODES_hostAlloc(n * b.nrows * b.nsens * m.neq * 8)
            if not hsp:
                raise SystemExit("bench.py: pinned host allocation failed")
            hsens = np.ctypeslib.as_array(C.cast(hsp, C.POINTER(C.c_double)), shape=(n, b.nrows, b.nsens, m.neq))
        chunk = 0                                                  # one launch; finished groups of sets stream back while it runs
        vals[:] = b.get_values(n)                                  # this step's device-generated sets, as host rows
        for _ in range(2):
            b.run_host(vals, res, hst, chunk=chunk, sens=hsens)
        launches1 = L.ODESBatch_getNumLaunches(b.h)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            b.run_host(vals, res, hst, chunk=chunk, sens=hsens)
        dt = time.perf_counter() - t0
        barrier()
        dt = max_over_ranks(dt)
        e2e_launches = L.ODESBatch_getNumLaunches(b.h) - launches1
        assert int((hst != 0).sum()) == failed
        # the host-buffer path must deliver the device-resident results (same seeded sets)
        probe = b.results(min(n, 4096))
        assert np.array_equal(probe, np.transpose(res[:probe.shape[2]], (1, 2, 0))), "host path and device-resident path disagree"
        if hsens is not None:
            assert np.array_equal(hsens[:256], b.sensitivities(256)), "host path and device-resident sensitivities disagree"
        e2e = {"value": world * n * args.steps / dt, "unit": UNIT, "h2d_bytes_per_step": n * nv * 8,
               "d2h_bytes_per_step": rows * n * 8 + n * 4 + (n * b.nrows * b.nsens * m.neq * 8 if b.nsens else 0),
               "ms_per_step": 1e3 * dt / args.steps, "chunk": chunk, "gpu_launches": e2e_launches}
        L.ODES_hostFree(hv); L.ODES_hostFree(hr); L.ODES_hostFree(hs); L.ODES_hostFree(hsp)

    # ---------------- CPU baseline on the same seeded sets (rank 0, N = 1 only) ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        kind = "reference" if H.have_ref() else "port"
        ns = args.cpu_sets or max(2048, min(4000 * cores, 400_000, n))
        if b.nsens:
            ns = args.cpu_sets or max(512, min(600 * cores, n))
        vals = b.get_values(ns)
        if b.nsens:                                                # the reference's vendored CVODES (oracle/_ref/libcvodes_ref.so)
            if not H.have_sens_ref():
                raise SystemExit("bench.py: oracle/_ref/libcvodes_ref.so is missing (build() makes it where /root/reference exists)")
            tp = H.tpoints(w["opt"], w["nout"])
            H.sens_ref_run(m, w["opt"], b, w["vary"], vals[:32], tp, 10000, w["rtol"], w["atol"], threads=cores, tag=args.workload + "_sens")
            t0 = time.perf_counter()
            ref = H.sens_ref_run(m, w["opt"], b, w["vary"], vals, tp, 10000, w["rtol"], w["atol"], threads=cores, tag=args.workload + "_sens")
            dt = time.perf_counter() - t0
            ref["stats"] = ref["stats"][:, :7]
            sens_same = bool(np.array_equal(b.sensitivities(ns), ref["sens"], equal_nan=True))
        else:
            cso = H.compile_c_model(m, args.workload)
            dts = []
            for _ in range(3):                                     # best of 3
                t0 = time.perf_counter()
                ref = H.oracle_run(m, w["opt"], w["vary"], vals, w["nout"], backend=H.REF if kind == "reference" else H.PORT, threads=cores, compiled_so=cso)
                dts.append(time.perf_counter() - t0)
            dt = min(dts)
            n1 = max(256, ns // (4 * cores))                       # the single-core figure
            t0 = time.perf_counter()
            H.oracle_run(m, w["opt"], w["vary"], vals[:n1], w["nout"], backend=H.REF if kind == "reference" else H.PORT, threads=1, compiled_so=cso)
            single_core = n1 / (time.perf_counter() - t0)
            sens_same = None
        gpu = b.results(ns)
        rep = H.parity_report(gpu, np.transpose(ref["out"][:, :, :m.neq], (1, 2, 0)), stats[:, :ns].T, ref["stats"], w["rtol"], w["atol"])
        cpu = {"value": ns / dt, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"first {ns} sets of this step's batch, {cores} threads, vendored {'CVODES 2.3.0 + emitted C rhs/Jacobian/sense_f' if b.nsens else 'CVODE 2.3.0 + emitted C rhs/Jacobian'}, gcc -O2 -ffp-contract=off",
               "cpu_model": cpu_model_name(), "timing": "best of 3, wall clock around the integration loop" if not b.nsens else "one run",
               "single_core": None if b.nsens else {"value": single_core, "unit": UNIT, "sets": n1},
               "failed_by_flag": status_histogram(ref["status"]),
               "sensitivities_bit_identical_on_sample": sens_same,
               "parity_on_sample": {"frac_within_10rtol": rep["frac_within"], "frac_same_step_sequence": rep["frac_same_path"], "max_ratio": rep["max_ratio"]},
               "oracle_mean_stats": dict(zip("nst nfe nsetups nje nni ncfn netf".split(), [float(x) for x in ref["stats"].mean(axis=0)]))}

    extra = {}
    if args.workload == "mapk" and not b.nsens and not args.no_other:
        if rank == 0 and world == 1:
            cores = os.cpu_count() or 1
            if not args.no_e2e:
                extra["e2e_entry"] = entry_leg(H, so, L, w, args.entry_sets, max(1, min(args.steps, 3)))
            extra["other_configs"] = [other_config_line(H, so, nm, 100_000, fp64_peak, cores, local) for nm in ("repressilator", "events", "huang", "goodwin")]
            extra["fmad_variant"] = fmad_variant(H, so, L, w, n, fp64_peak, local, cores)
        if world > 1:
            # config 5 on all GPUs of the job (every rank takes part)
            line5 = scaled_config_line(H, so, "huang", 100_000, local, world, rank, barrier, max_over_ranks)
            if rank == 0:
                extra["other_configs"] = [line5]
    barrier()
    if rank == 0:
        Ff, FJ = count_ops(b.source())
        # the work counters of the ORACLE (the reference's CVODE on the CPU sample) where that leg ran, else the device's
        flop_stats = cpu["oracle_mean_stats"] if cpu else mean_stats
        F = algorithmic_flops(m.neq, Ff, FJ, flop_stats)
        if b.nsens:
            # the simultaneous corrector adds, per right-hand-side evaluation, one Jacobian and per sensitivity a sparse
            # matrix-vector product (2 nnz; the parametric term is neglected), per Newton iteration one more solve per
            # sensitivity, per step the history operations of every sensitivity vector
            nnz, nS, nn, qb = len(m.jacobian_structure()), b.nsens, m.neq, 4.0
            F += mean_stats["nfe"] * (FJ + nS * 2 * nnz) + mean_stats["nni"] * nS * (2 * nn * nn + 6 * nn) + mean_stats["nst"] * nS * nn * (qb * (qb + 1) / 2 + 2 * (qb + 1) + 10)
        achieved = F * n / (kernel_ms * 1e-3) / 1e12
        out_bytes = 8 * b.nrows * b.nstore + 8 * nv + 8 * b.nrows * b.nsens * m.neq
        traffic, traffic_src = None, None
        for prof_name in ("r02_ncu_metrics.json", "r01_ncu_metrics.json"):        # one ncu --set full capture of this kernel (profiles/)
            prof = os.path.join(ROOT, "profiles", prof_name)
            if traffic is None and os.path.exists(prof) and args.workload == "mapk" and not b.nsens:
                pj = json.load(open(prof))
                if pj.get("sets") and pj.get("dram_bytes") is not None:
                    traffic = pj["dram_bytes"] / pj["sets"] * n    # per launch; measured at this size when the capture's sets == n
                    traffic_src = f"profiles/{prof_name}: dram__bytes_read.sum + dram__bytes_write.sum of one launch of {pj['sets']} sets" + ("" if pj["sets"] == n else f", scaled to {n}")
        info = b.kernel_info()
        workload_text = ("MAPK.xml parameter scan (BASELINE configs[1]): 21 kinetic constants x 2^U(-1,1), rtol 1e-6 atol 1e-9, t=0..1000, 100 output points, 8 species stored"
                         if args.workload == "mapk" and not b.nsens else
                         f"{args.workload} + {b.nsens} forward sensitivities ({args.sens}; SURVEY 8f rank 4, not the headline): CVODES simultaneous corrector, {nv} varied values, rtol 1e-6 atol 1e-9, t=0..{w['end']:g}, {w['nout']} output points, {m.neq} variables and {b.nsens} x {m.neq} sensitivities stored"
                         if b.nsens else
                         f"{args.workload} (BASELINE parity configuration, not the headline): {nv} varied values, rtol 1e-6 atol 1e-9, t=0..{w['end']:g}, {w['nout']} output points, {m.neq} variables stored")
        line = {"metric": METRIC if args.workload == "mapk" and not b.nsens else f"trajectories/sec ({args.workload}{' + %d sensitivities' % b.nsens if b.nsens else ''}, rtol 1e-6)", "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_text,
                           "sets_per_gpu": n, "nvary": nv, "seed": SEED, "sharding": f"range x{world}, no collective", "host_affinity": numa,
                           "l2": f"per step {1e-9 * n * (8 * b.nrows * b.nstore + 8 * nv):.1f} GB of outputs + inputs, far larger than the 126 MB L2; no explicit flush"},
                "failed_sets": failed, "mean_stats": mean_stats, "clocks": clocks, "e2e": e2e, "gpu_launches": launches,
                "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak if fp64_peak else None,
                             "traffic": traffic, "traffic_source": traffic_src, "flop_counters": "oracle (reference CVODE on the CPU sample)" if cpu else "device", "peak_source": "measured live: ODES_measureFP64Peak (FMA chain micro-benchmark); MEASURED_PEAKS.json holds no FP64 figure",
                             "flop_per_trajectory": F, "flop_model": {"F_f": Ff, "F_J": FJ, "qbar": 4.0, "formula": "SURVEY 8(d)"},
                             "kernel": "odes_bdf_kernel", "kernel_ms": kernel_ms, "kernel_info": info,
                             "hbm_view": {"algorithmic_bytes_per_trajectory": out_bytes, "achieved_gbs": out_bytes * n / (kernel_ms * 1e-3) / 1e9,
                                          "peak_gbs": json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0}},
                "cpu_baseline": cpu}
        line.update(extra)
        ceil = os.path.join(ROOT, "profiles", "r02_d2h_ceiling.json")
        if e2e and os.path.exists(ceil):
            # what the host of this pool's boxes absorbs (plain pinned cudaMemcpyAsync, tools/micro/d2h_ceiling.cu)
            cj = {r["gpus"]: r["aggregate_gbs"] for r in json.load(open(ceil))["runs"] if r["direction"] == "d2h" and r["shape"] == "106MB groups"}
            if world in cj:
                got = world * e2e["d2h_bytes_per_step"] / (e2e["ms_per_step"] * 1e-3) / 1e9
                e2e["d2h_gbs"] = got
                e2e["bound"] = f"host D2H {got:.1f} GB/s of {cj[world]:.1f} GB/s ceiling at {world} GPU(s) (profiles/r02_d2h_ceiling.json)" if got > 0.85 * cj[world] else f"kernel (D2H {got:.1f} GB/s of {cj[world]:.1f} GB/s ceiling)"
        rec = os.path.join(ROOT, "profiles", "MEASURED_FP64_PEAK.json")
        if os.path.exists(rec):
            line["roofline"]["peak_recorded"] = json.load(open(rec))
        print(json.dumps(line), flush=True)
    b.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
