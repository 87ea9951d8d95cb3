#!/usr/bin/env python
"""bench.py — particle-timesteps/s of OpenRBC's per-timestep force/integrate loop on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload rbc|sphere|patch:<n_lipids>] [--impl ours|reference]

One "step" is one iteration of the reference's main MD loop (src/openrbc.cpp:189-256, default LANGEVIN build) over the
whole system: [every 2nd step: Voronoi/cell rebuild] -> pair forces -> bonded forces -> Langevin integrator, with the
stray-lipid cleanup every 60 steps.  Default workload = BASELINE.json configs[1]: the full red blood cell
(`-i trimesh -m rbc`, 3 205 506 particles, 188 549 Voronoi cells), initial state produced by the reference's own host-side
initialisation (init_rbc + VoronoiDiagram::init + 100 minimisation steps; tools/make_states.py), which the north star
leaves in place.

Printed JSON (one line, rank 0):
  value     device-resident throughput of the fused whole-loop entry point orbc_run_langevin (state already in HBM)
  e2e       same loop driven call by call through the reference-facing C ABI from HOST buffers: upload of the containers
            from pinned memory, K x (rebuild / compute_pairwise_fused / compute_bonded / integrate + status read-back),
            temperature every 100 steps, and the download a save_frame needs at the end — all inside the timed region
  roofline  dominant kernel (lipid pair forces): algorithmic bytes per launch / mean launch duration (CUDA events on the
            launching stream, recorded live in the timed region) against the measured HBM peak
  cpu_baseline  the unmodified reference (oracle/_ref/libref_fast.so, as-shipped flags) on this box's host cores

`--impl reference` times only the reference's own OpenMP implementation of the same loop on the same state.
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "particle-timesteps/sec"
UNIT = "particle-steps/s"
B_ALG_STEP = 96.0            # algorithmic bytes per particle-timestep: read + write x, v, n, o as 3 x fp32
B_ALG_PAIR = 48.0            # pair-force kernel: read x, n (24 B), write f, t (24 B) per particle and launch
FREQ_CLEANUP, FREQ_DISPLAY = 60, 100      # runtime_parameter.h:59-60
HBM_FALLBACK_GBS = 6650.0    # /opt/skills/guides/B200_PROFILING.md fallback


# ---------------------------------------------------------------------------------------------------------------------
# workloads
# ---------------------------------------------------------------------------------------------------------------------
def state_path(workload):
    return os.path.join(ROOT, "data", "_gen", workload.replace(":", "_") + ".npz")


def ensure_state(workload):
    """Input state of the workload.  rbc / sphere come from the reference's host-side initialisation (code the north star
    leaves in place), run once through tools/make_states.py and cached; patch:<n> is the synthetic bilayer of §8(d)."""
    path = state_path(workload)
    if os.path.exists(path):
        return path
    t0 = time.time()
    if workload.startswith("patch:"):
        from openrbc_b200.synthetic import flat_patch_state
        st = flat_patch_state(int(float(workload.split(":")[1])))
        os.makedirs(os.path.dirname(path), exist_ok=True)
        np.savez(path, **st)
    else:
        kind = {"rbc": ["rbc", "--opt", "100"], "sphere": ["sphere", "--radius", "100", "--opt", "20"]}[workload]
        # torchrun exports OMP_NUM_THREADS=1 to its workers; the reference's initialisation should use the whole host
        env = dict(os.environ, OMP_PROC_BIND="close", OMP_PLACES="cores", OMP_NUM_THREADS=str(host_threads()))
        out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "make_states.py"), *kind, "--out", path],
                             env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if out.returncode != 0 or not os.path.exists(path):
            raise SystemExit("could not generate the %s state (needs oracle/_ref/libref_fast.so built by `python __graft_entry__.py build`):\n%s"
                             % (workload, out.stdout[-2000:]))
    print(f"[bench] generated {path} in {time.time() - t0:.1f}s", file=sys.stderr, flush=True)
    return path


def load_state(workload):
    return dict(np.load(ensure_state(workload)))


def data_label(workload):
    if workload == "rbc":
        return "real mesh: example-large/rbc.*.txt shipped with the reference, membrane and cytoskeleton placed by the reference's own init_rbc (seed 0xBAD5EED), 100 minimisation steps; no external dataset"
    if workload == "sphere":
        return "synthetic: the reference's own random lipid sphere (init_random_sphere, R = 100)"
    return "synthetic: flat bilayer patch (openrbc_b200/synthetic.py, SURVEY.md 8d config 4)"


def workload_name(workload, st):
    n = len(st["lx"]) + len(st["px"])
    base = {"rbc": "full RBC, -i trimesh -m rbc (example-large), Langevin", "sphere": "lipid sphere R=100, -i lipid, Langevin"}.get(
        workload, f"synthetic flat bilayer {workload}, Langevin")
    return f"{base}: {n} particles, {len(st['centroids'])} Voronoi cells, {len(st['bonds'])} bonds"


# ---------------------------------------------------------------------------------------------------------------------
# reference arm: the unmodified reference on the host cores
# ---------------------------------------------------------------------------------------------------------------------
def host_threads():
    """Host threads this process may use (cgroup / affinity aware)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def reference_child(args):
    """One measurement of the unmodified reference at a fixed OpenMP thread count (its own process: the reference keeps
    function-local statics sized by omp_get_max_threads()).  Prints the JSON line, then silences stdout so that the
    reference's at-exit timer report (timer.h:55) cannot follow it."""
    from oracle import ref as refmod
    st = load_state(args.workload)
    n = len(st["lx"]) + len(st["px"])
    threads = args.threads
    r = refmod.Ref("fast", threads=threads, args=["-i", "lipid"])   # the state is loaded below; "lipid" only satisfies the CLI check
    r.load_state(st)
    r.set_param("kBT", 0.22)
    budget = args.ref_budget
    t0 = time.time()
    r.run_langevin(args.warmup, cleanup=True)
    per = max((time.time() - t0) / max(args.warmup, 1), 1e-3)
    steps = int(max(2, min(args.steps, (budget - (time.time() - t0)) / per)))
    sec = r.run_langevin(steps, cleanup=True)
    value = n * steps / sec
    sample = f"{steps} MD steps of the whole system after {args.warmup} warm-up steps ({sec:.2f} s)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sec / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": data_label(args.workload),
        "config": {"workload": workload_name(args.workload, st), "impl": "unmodified reference headers, g++ -O3 -march=x86-64-v3 -mtune=generic -ffast-math -mrecip -fopenmp (oracle/Makefile FAST: AVX2/FMA ISA level instead of the "
                                                                       "as-shipped -march=native, because the library is built on another machine than it runs on), compute_pairwise_fused path",
                   "omp_threads": threads, "omp_binding": "OMP_PROC_BIND=%s OMP_PLACES=%s" % (os.environ.get("OMP_PROC_BIND", "unset"), os.environ.get("OMP_PLACES", "unset")),
                   "pair_timer_s_per_step": r.timer("compute_pairwise_fused") / max(steps + args.warmup, 1)},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    sys.stdout.flush()
    os.dup2(os.open(os.devnull, os.O_WRONLY), 1)


def reference_line(args, steps, warmup, budget):
    """Runs reference_child with all host threads; if that process dies (the reference was written for tens of threads, not
    hundreds) the thread count is halved until a run completes.  Returns (json dict or None, list of attempts)."""
    from oracle import ref as refmod
    if not refmod.available("fast"):
        return {"impl": "reference", "unavailable": "oracle/_ref/libref_fast.so is not built (python __graft_entry__.py build where /root/reference exists)"}, []
    ensure_state(args.workload)
    cand, t = [], args.threads or host_threads()
    while t >= 1 and len(cand) < 6:
        cand.append(t)
        if args.threads:
            break
        t //= 2
    attempts = []
    for t in cand:
        env = dict(os.environ, OMP_NUM_THREADS=str(t))
        env.setdefault("OMP_PROC_BIND", "close"); env.setdefault("OMP_PLACES", "cores")
        for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
            env.pop(k, None)
        cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference-child", "--workload", args.workload, "--steps", str(steps),
               "--warmup", str(warmup), "--ref-budget", str(budget), "--threads", str(t), "--gpus", str(args.gpus)]
        try:
            out = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=budget * 3 + 300)
            for ln in reversed(out.stdout.strip().splitlines()):
                if ln.startswith("{"):
                    d = json.loads(ln)
                    d["config"]["thread_attempts"] = attempts + [f"{t}: ok"]
                    return d, attempts
            attempts.append(f"{t}: exit {out.returncode} {out.stderr.strip().splitlines()[-1][:120] if out.stderr.strip() else ''}")
        except Exception as e:  # noqa: BLE001
            attempts.append(f"{t}: {type(e).__name__}")
    return None, attempts


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    d, attempts = reference_line(args, args.steps, args.warmup, args.ref_budget)
    if d is None:
        d = {"impl": "reference", "unavailable": "the reference process failed at every thread count tried: " + "; ".join(attempts)}
    print(json.dumps(d), flush=True)


def cpu_baseline_subprocess(args):
    """Reference timed on this box's host cores in its own process (one OpenMP runtime, one thread count per process)."""
    d, attempts = reference_line(args, args.cpu_steps, 2, 40)
    if d and "cpu_baseline" in d:
        cb = d["cpu_baseline"]
        cb["omp"] = d["config"]["omp_binding"]
        return cb
    why = d.get("unavailable") if d else "failed: " + "; ".join(attempts)
    return {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": why}


# ---------------------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------------------------------------
class Clocks:
    """SM clock / power / throttle reasons sampled every 20 ms through NVML while the timed region runs."""

    def __init__(self, index):
        self.rows, self.stop, self.index = [], False, index
        self.th = threading.Thread(target=self._run, daemon=True)
        self.max_sm = None
        self.err = None

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else self.index
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.max_sm = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            while not self.stop:
                self.rows.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetPowerUsage(h) / 1e3,
                                  nv.nvmlDeviceGetCurrentClocksEventReasons(h)))
                time.sleep(0.02)
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)

    def __enter__(self):
        self.th.start()
        time.sleep(0.05)
        return self

    def __exit__(self, *a):
        self.stop = True
        self.th.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": ["nvml unavailable: %s" % self.err]}
        sm = sorted(r[0] for r in self.rows)
        bits = 0
        for r in self.rows:
            bits |= r[2]
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap", 0x80: "hw_power_brake_slowdown"}
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_sm, "reasons": [n for b, n in names.items() if bits & b], "samples": len(sm),
                "power_w_max": max(r[1] for r in self.rows)}


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------
def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            for k in ("hbm_gbs", "hbm_gbs_sustained"):
                if k in d:
                    return float(d[k]), "measured (MEASURED_PEAKS.json %s)" % k
        except Exception:  # noqa: BLE001
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic(mix):
    """DRAM bytes per launch of the dominant kernel -- the lipid pair evaluation, which is one of three kernels (search, record,
    walk) by the gate's decision -- from the committed `ncu --set full` captures (profiles/ncu_traffic.json, written by hand from
    profiles/r02_*_ncu.txt: dram__bytes_read.sum + dram__bytes_write.sum of one launch each), weighted by how often each ran in
    the timed region (`mix` = {"search": n, "record": n, "walk": n}); None when absent."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        tot = sum(mix.values())
        if not tot:
            return None, None
        t = sum(n * (float(d[k]["dram_bytes_read"]) + float(d[k]["dram_bytes_write"])) for k, n in mix.items() if n) / tot
        return t, "; ".join("%s x%d: %s" % (k, n, d[k].get("source")) for k, n in mix.items() if n)
    except Exception:  # noqa: BLE001
        return None, None


def alu_roofline(prof, counts, world):
    """FP32 roofline of the pair kernels (SURVEY.md 8d 'Algorithmic flops'): flops per force evaluation / mean device time of the
    two pair kernels (CUDA events, live) against the FFMA peak measured on this pool's B200 (tools/microbench/fp32_peak.cu ->
    profiles/r02_fp32_peak.json).  executed = what the one-sided GPU evaluation does (every in-cutoff pair from both sides);
    algorithmic = the reference's Newton-on count (each pair once).  98 flop per evaluated pair + 8 per distance test."""
    try:
        pk = json.load(open(os.path.join(ROOT, "profiles", "r02_fp32_peak.json")))
    except Exception:  # noqa: BLE001
        pk = {}
    peak = pk.get("ffma_tflops")
    if not counts or world != 1:
        return {"bound": "fp32", "peak": peak, "unit": "TFLOP/s", "note": "pair statistics are taken on the single-GPU run only"}
    t_ll, h_ll, t_pl, h_pl, t_pp, h_pp = counts                      # one-sided tests / hits of LL and PP; PL from the protein side
    (ms_l, n_l), (ms_p, n_p) = prof["pair_lipid"], prof["pair_protein"]
    sec = (ms_l / max(n_l, 1) + ms_p / max(n_p, 1)) * 1e-3          # one evaluation of both kernels, averaged over building and list-walking steps
    executed = (h_ll + 2 * h_pl + h_pp) * 98.0 + (t_ll + t_pl + t_pp) * 8.0
    algorithmic = (h_ll / 2 + h_pl + h_pp / 2) * 106.0 + (t_ll / 2 + t_pl + t_pp / 2) * 8.0
    out = {"bound": "fp32", "unit": "TFLOP/s", "peak": peak, "peak_source": "measured: tools/microbench/fp32_peak.cu on this pool's B200 (profiles/r02_fp32_peak.json); issue ceiling %.3g warp instructions/s" % pk.get("mixed_warp_inst_per_s", float("nan")),
           "kernels": "lipid + protein pair kernels of one force evaluation", "seconds_per_evaluation": sec,
           "pairs": {"ll_tests": t_ll, "ll_hits_one_sided": h_ll, "pl_tests": t_pl, "pl_hits": h_pl, "pp_tests": t_pp, "pp_hits_one_sided": h_pp},
           "gflop_executed": executed / 1e9, "gflop_algorithmic": algorithmic / 1e9,
           "achieved_executed": executed / sec / 1e12 if sec else None, "achieved_algorithmic": algorithmic / sec / 1e12 if sec else None,
           "ncu": "profiles/r02_search_kernel_ncu.txt (sm__throughput 63 %, fma pipe 34 %, alu pipe 38 %, issue active 65 %), profiles/r02_record_kernel_ncu.txt (issue 68 %), "
                  "profiles/r02_list_walker_ncu.txt (issue 69 %, fma pipe 43 %), profiles/r02_prot_search_ncu.txt (issue 50 %), profiles/r02_prot_walker_ncu.txt (issue 18 %)"}
    if peak and sec:
        out["frac_executed"] = out["achieved_executed"] / peak; out["frac_algorithmic"] = out["achieved_algorithmic"] / peak
    return out


def run_chunked(sim, n_steps, nh=False, frames=0, sink=None):
    """n_steps of the main loop through the fused entry point, with delete_lipid at multiples of freq_cleanup (openrbc.cpp:201).
    nh: the Nose-Hoover branch (orbc_run_nh).  frames: a trajectory frame (dump_field 7: x, n, affiliation, ids) every `frames` steps,
    assembled on the device and copied out while the run continues (openrbc.cpp:248-250); sink(frame bytes) receives it."""
    done, pending = 0, False
    while done < n_steps:
        if sim.nstep % FREQ_CLEANUP == 0:
            sim.delete_lipid(sim.stray_tolerance)
        k = min(n_steps - done, FREQ_CLEANUP - sim.nstep % FREQ_CLEANUP)
        if frames:
            k = min(k, frames - sim.nstep % frames)
        (sim.run_nh if nh else sim.run_langevin)(k)
        done += k
        if frames and sim.nstep % frames == 0:
            if pending:
                sink(sim.save_frame_end())
            sim.save_frame_begin(7); pending = True
    if pending:
        sink(sim.save_frame_end())


def run_ours(args):
    import torch
    import torch.distributed as dist
    import openrbc_b200 as orbc

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — openrbc_b200 has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        ensure_state(args.workload)
    if world > 1:
        dist.barrier()
    st = load_state(args.workload)
    n_total = len(st["lx"]) + len(st["px"])

    def connect(s):
        """Exchange the ranks' connection blobs (CUDA IPC handles) once; the data path never touches torch.distributed again."""
        blobs = [None] * world
        dist.all_gather_object(blobs, s.mg_export())
        s.mg_connect(blobs)

    def make_sim(state):
        s = orbc.Simulation(state, kBT=0.22, device=local, rank=rank, world=world)
        s.stray_tolerance = 2.5
        for kv in args.opt:
            k, v = kv.split("=")
            s.set_option(k, float(v))
        if args.cv:
            s.set_volume_constraint(True, 3.15, 0.05)              # inside orbc_run_langevin; the per-call e2e loop calls constrain_volume itself
        if world > 1:
            connect(s)
        return s

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident leg ------------------------------------------------------------------------------------------
    sim = make_sim(st)
    frame_bytes = [0]
    devnull = open(os.devnull, "wb")

    def sink(fr):
        devnull.write(memoryview(fr)); frame_bytes[0] += fr.nbytes
    if args.frames and world > 1:
        raise SystemExit("bench.py: --frames is a single-GPU option (a rank's frame holds its own slots only)")
    if args.nh:
        sim.zeta = 0.0
    if args.frames:                      # the two pinned frame buffers are allocated on first use: not part of the steady state
        sim.save_frame_begin(7); sim.save_frame_begin(7); sim.save_frame_end(); sim.save_frame_end()
    run_chunked(sim, args.warmup, args.nh)
    sim.synchronize()
    nl_before = sim.dump("nl_stats").tolist()
    sim.profile_enable(True)
    l0 = sim.launch_count()
    barrier()
    with Clocks(local) as clk:
        t_dev0 = time.perf_counter()
        sim.event_record(0)
        run_chunked(sim, args.steps, args.nh, args.frames, sink)
        sim.event_record(1)
        sim.synchronize()
        t_dev1 = time.perf_counter()
        barrier()
    ms = sim.event_elapsed_ms(0, 1)
    if args.frames:                      # the frames leave on a second stream: the job is done when the last one has arrived
        ms = max(ms, (t_dev1 - t_dev0) * 1e3)
    launches = sim.launch_count() - l0
    prof = {k: sim.profile_read(k) for k in orbc.engine.PROF}
    sim.profile_enable(False)
    n_now = sim.size(0) + sim.size(1)
    nl_stats = sim.dump("nl_stats").tolist()
    nl_timed = [int(a) - int(b) for a, b in zip(nl_stats, nl_before)]      # [recorded, walked, -, searched] inside the timed region
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # N>1: ONE cell decomposed over the ranks (strong scaling): the job is still n_total particles x steps
    value = n_total * args.steps / (ms * 1e-3)
    temperature = sim.compute_temperature()
    if world > 1:
        t = torch.tensor([temperature, float(n_now)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)                                       # every rank returns its additive share of sum(m v^2) / 3N
        temperature = float(t[0].item()); n_now = n_total

    # ---- pair statistics of the current state for the ALU roofline (untimed): the cross-check kernels count tests and hits ------
    pair_counts = None
    if world == 1:
        c0 = sim.dump("counters").astype(np.int64)
        sim.set_option("pair_impl", 1); sim.clear_force(); sim.compute_pairwise_fused(); sim.synchronize()
        pair_counts = (sim.dump("counters").astype(np.int64) - c0)[1:7].tolist()
        sim.set_option("pair_impl", 2); sim.clear_force()

    # ---- end-to-end leg: host buffers through the per-call C ABI ---------------------------------------------------------
    # the host's containers in pinned memory: the eight vector arrays and the index arrays (ids, bonds, Voronoi state) alike —
    # 15 MB of pageable index arrays cost as much PCIe time as the 154 MB of pinned vectors
    pinned = {}
    for k in ("lx", "lv", "ln", "lo", "px", "pv", "pn", "po", "ptype", "ptag", "bonds", "centroids", "cs_l", "cs_p"):
        a = np.ascontiguousarray(st[k], np.float32 if st[k].dtype.kind == "f" else np.int32)
        tns = torch.empty(a.shape, dtype=torch.float32 if a.dtype == np.float32 else torch.int32, pin_memory=True)
        tns.numpy()[...] = a
        pinned[k] = tns
    host = dict(st)
    host.update({k: v.numpy() for k, v in pinned.items()})
    out_l = {f: torch.empty((len(st["lx"]), 3), dtype=torch.float32, pin_memory=True) for f in "xn"}
    out_p = {f: torch.empty((len(st["px"]), 3), dtype=torch.float32, pin_memory=True) for f in "xn"}
    aff_l = torch.empty(len(st["lx"]), dtype=torch.int32, pin_memory=True)
    aff_p = torch.empty(len(st["px"]), dtype=torch.int32, pin_memory=True)
    # the e2e job re-uses the context (device allocations and, on N>1, the peer mappings are set-up, not per-job work)
    e2e_sim = sim
    barrier()
    # N>1: every rank uploads the rows of its own cells only (orbc_upload_range); the halo copies come from their owners over NVLink
    e2e_sim.upload(host, owned_only=world > 1)
    if world > 1:
        e2e_sim.mg_export()
    e2e_sim.nstep = 0
    for _ in range(min(args.warmup, 4)):
        e2e_sim.step_langevin_checked()
    e2e_sim.nstep = 0
    barrier()
    e2e_sim.event_record(2)
    t_wall = time.perf_counter()
    e2e_sim.upload(host, owned_only=world > 1)
    t_up0 = time.perf_counter()
    if world > 1:
        e2e_sim.mg_export()              # owned ranges + halo masks of the fresh state (device work, no new allocations); collective
    t_up = time.perf_counter()
    trace = bool(os.environ.get("ORBC_BENCH_TRACE"))
    if trace:
        print(f"[bench rank {rank}] e2e upload {(t_up0 - t_wall) * 1e3:.1f} ms {e2e_sim.last_upload_ms} (lipids, proteins, bonds, voronoi) + export {(t_up - t_up0) * 1e3:.1f} ms", file=sys.stderr, flush=True)
    slow = []                            # (ms, what) of the slowest host-side calls, for the phase breakdown
    h2d = e2e_sim.last_upload_bytes      # this rank's rows of x, v, n, o + ids, bonds and the Voronoi state (the job total is the sum over ranks)
    d2h = 0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        if e2e_sim.nstep % FREQ_CLEANUP == 0:
            e2e_sim.delete_lipid(e2e_sim.stray_tolerance); d2h += 8
            slow.append(((time.perf_counter() - t0) * 1e3, "delete_lipid@%d" % e2e_sim.nstep)); t0 = time.perf_counter()
        if args.nh:
            e2e_sim.step_nh_checked(); d2h += 32
        elif args.cv:
            e2e_sim.step_langevin_cv_checked(3.15, 0.05); d2h += 20
        else:
            e2e_sim.step_langevin_checked(); d2h += 16
        slow.append(((time.perf_counter() - t0) * 1e3, "step@%d" % (e2e_sim.nstep - 1)))
        if trace and slow[-1][0] > 20.0:
            print(f"[bench rank {rank}] slow call {slow[-1]}", file=sys.stderr, flush=True)
        if e2e_sim.nstep % FREQ_DISPLAY == 0:
            e2e_sim.compute_temperature(); d2h += 8
    t_steps = time.perf_counter()
    fl = e2e_sim.download_into(0, x=out_l["x"].numpy(), n=out_l["n"].numpy(), affiliation=aff_l.numpy())
    fp = e2e_sim.download_into(1, x=out_p["x"].numpy(), n=out_p["n"].numpy(), affiliation=aff_p.numpy())
    d2h += fl + fp
    e2e_sim.event_record(3)
    e2e_sim.synchronize()
    wall_ms = (time.perf_counter() - t_wall) * 1e3
    slow.sort(reverse=True)
    phases = {"upload_ms": (t_up - t_wall) * 1e3, "steps_ms": (t_steps - t_up) * 1e3, "download_ms": wall_ms - (t_steps - t_wall) * 1e3,
              "slowest_calls": [[round(a, 2), b] for a, b in slow[:4]], "median_step_ms": round(slow[len(slow) // 2][0], 3)}
    if os.environ.get("ORBC_BENCH_TRACE"):
        print(f"[bench rank {rank}] e2e phases {phases}", file=sys.stderr, flush=True)
        # a few more (untimed) steps with a host synchronisation behind every call: which call is the slow one?
        acc = {}
        for _ in range(8):
            calls = [("rebuild", e2e_sim.rebuild)] if e2e_sim.nstep % e2e_sim.freq_voronoi == 0 else []
            calls += [("pairwise", e2e_sim.compute_pairwise_fused), ("bonded", e2e_sim.compute_bonded), ("langevin", e2e_sim.verlet_langevin)]
            for name, fn in calls:
                t0 = time.perf_counter(); fn(); t1 = time.perf_counter(); e2e_sim.synchronize(); t2 = time.perf_counter()
                a = acc.setdefault(name, [0.0, 0.0, 0]); a[0] += (t1 - t0) * 1e3; a[1] += (t2 - t1) * 1e3; a[2] += 1
            e2e_sim.nstep += 1
        print(f"[bench rank {rank}] per call (issue ms, wait ms): " + ", ".join(f"{k} {v[0] / v[2]:.3f}/{v[1] / v[2]:.3f}" for k, v in acc.items()), file=sys.stderr, flush=True)
    barrier()
    e2e_ms = max(e2e_sim.event_elapsed_ms(2, 3), wall_ms)
    if world > 1:
        t = torch.tensor([e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = n_total * args.steps / (e2e_ms * 1e-3)
    barrier()
    e2e_sim.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = hbm_peak()
    alu = alu_roofline(prof, pair_counts, world)
    pl_ms, pl_cnt = prof["pair_lipid"]
    n_l = len(st["lx"]) / world          # lipids one launch of the kernel covers (this rank's share on N>1)
    achieved = (B_ALG_PAIR * n_l / (pl_ms / pl_cnt * 1e-3) / 1e9) if pl_cnt else None
    # which of the three lipid kernels ran in the timed region (hit-list statistics of the timed steps; without lists: all searches)
    ll_mix = {"search": nl_timed[3] if (nl_timed[0] + nl_timed[1] + nl_timed[3]) else int(pl_cnt), "record": nl_timed[0], "walk": nl_timed[1]}
    shares = {k: round(v[0] / ms, 4) for k, v in prof.items()}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": data_label(args.workload),
        "config": {"workload": workload_name(args.workload, st), "l2": "inputs larger than L2 (state %.0f MB resident in HBM, no flush needed)" % (n_total * 6 * 16 / 1e6),
                   "multi_gpu": "single GPU" if world == 1 else (f"one cell decomposed over {world} ranks: contiguous ranges of Morton-ordered Voronoi cells, halo push + migration "
                                                                "by peer stores over NVLink, epoch-flag barriers"),
                   "integrator": ("Nose-Hoover (fused initial / final kernels) kBT=0.22 dt=0.01" if args.nh else "verlet_langevin kBT=0.22 dt=0.01") + ", rebuild every 2 steps, Morton sort every 24, cleanup every 60"
                                 + (", constrain_volume(3.15, 0.05) every step" if args.cv else "")
                                 + ((", a %.0f MB trajectory frame every %d steps assembled on the device and copied out while the run continues (%d frames in the timed region)"
                                     % (frame_bytes[0] / max(1, args.steps // args.frames) / 1e6, args.frames, args.steps // args.frames)) if args.frames else ""),
                   "particles_at_end": n_now, "temperature_at_end": temperature, "options": args.opt,
                   "hit_lists": {"evaluations_that_recorded": nl_timed[0], "evaluations_that_walked": nl_timed[1], "evaluations_that_searched": nl_timed[3], "overflow": nl_stats[2], "of": "the timed steps"},
                   "hbm_roofline_frac_step": value / world * B_ALG_STEP / 1e9 / peak, "device_time_shares": shares},
        "roofline": {"bound": "hbm", "kernel": "lipid side of compute_pairwise_fused: k_pair_ll_r<.,4,false> (search), k_pair_ll_r<.,4,true> (search + record the hit lists) or "
                               "k_pair_ll_list (walk them), by the device's gate; mean over the launches of the timed region", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": (ncu_traffic(ll_mix)[0] if world == 1 else None),
                     "traffic_source": ncu_traffic(ll_mix)[1], "launch_mix": ll_mix, "peak_source": peak_src,
                     "launch_ms": pl_ms / pl_cnt if pl_cnt else None, "launches_timed": pl_cnt,
                     "algorithmic_bytes_per_launch": B_ALG_PAIR * n_l,
                     "note": "pair forces are FP32-ALU bound (~1.6-3 kFLOP per 48 B), see DESIGN.md; HBM fraction reported as the contract asks"},
        "roofline_alu": alu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
                "ms_per_step": e2e_ms / args.steps, "phases": phases, "what": "upload from pinned host containers + K per-call steps with status read-back + frame download, all timed"},
        "gpu_launches": int(launches),
        "clocks": clk.summary(),
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_subprocess(args)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=240)
    ap.add_argument("--warmup", type=int, default=24)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-child"])
    ap.add_argument("--workload", default="rbc")
    ap.add_argument("--threads", type=int, default=0, help="reference arm: OpenMP threads (0 = all host cores)")
    ap.add_argument("--ref-budget", type=float, default=150.0, help="reference arm: wall-clock bound in seconds")
    ap.add_argument("--cpu-steps", type=int, default=40, help="steps of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--opt", action="append", default=[], help="library option name=value (orbc_set_option), e.g. --opt nl_skin=0.2")
    ap.add_argument("--cv", action="store_true", help="BASELINE.json configs[2]: constrain_volume(3.15, 0.05) at the place of openrbc.cpp:229 in every step")
    ap.add_argument("--nh", action="store_true", help="BASELINE.json configs[4]: the Nose-Hoover branch of the loop instead of Langevin")
    ap.add_argument("--frames", type=int, default=0, help="configs[4]: a trajectory frame every N steps of the device-resident leg (single GPU)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"], help="label of the run (weak: the workload was sized with the number of GPUs, e.g. patch:<N x 1.05e6>)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.impl == "reference-child":
        reference_child(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
