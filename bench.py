#!/usr/bin/env python
"""Headline benchmark of the ANN bias-force hot path (contract in the task statement, §④).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c5] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic replicas: every replica's
coordinates -> MLP forward -> umbrella energy -> analytic backward -> forces.  The default workload
is BASELINE.json configs[4] sharded the way it is quoted: 4500-512-512-8 encoder on 1500 atoms,
1024 replicas PER GPU (8192 on 8 GPUs) — weak scaling, no collective on the force path.  Other
BASELINE shapes are selectable with --workload (c2 single replica on OpenMM buffers, c3, c4).

Prints ONE JSON line (rank 0).  `value` is whole-job evaluations/s with inputs resident in HBM (the timed steps are
replayed from a CUDA graph over rotating buffer sets larger than L2; device time, max over ranks); `e2e` is the same metric
through the host-buffer C-ABI entry (H2D + eval + D2H every step) — for the single-replica workload, through the C++ plugin
layer on device-resident OpenMM buffers; `roofline` comes from a second, event-instrumented pass (never graph-replayed);
`cpu_baseline` is the reference's own CPU kernel on the host cores.  The default 1-GPU line also carries short runs of the
other BASELINE shapes (`other_workloads`: c2 with the legacy-restatement comparison, c3, c4).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from ann_force_b200 import synthetic  # noqa: E402

METRIC = "ANN bias force evals/sec (replicas x steps)"
UNIT = "evals/s"
L2_BYTES = 126 * 1024 * 1024
TIMESTEP_FS = 2.0          # ns/day = steps/s per replica * dt * 86400 s/day
MIN_TIMED_S = 0.3          # device time measured per run, whatever --steps is (windows of exactly --steps steps)
MIN_WINDOWS, MAX_WINDOWS = 5, 4000


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"],
                    bf16_tflops_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# --------------------------------------------------------------------------------------------------
# clocks during the timed region
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons through NVML.  `sample()` is only ever called from the main thread BETWEEN timed
    windows — after one window's stop event and before the next one's start event have been enqueued — while the GPU
    is still working through the queue of earlier windows: the clocks are those under load, and no NVML call (or any
    other host work) sits inside an event pair.  There is no polling thread."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x100: "display_clock_setting"}

    def __init__(self, device_index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def sample(self):
        if self.nvml is None:
            return
        try:
            self.samples.append(float(self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM)))
            mask = int(self.nvml.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
            for bit, name in self.REASONS.items():
                if mask & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def summary(self):
        if not self.samples:
            return dict(sm_mhz=None, sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons), samples=0)
        return dict(sm_mhz=float(np.median(self.samples)), sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons),
                    samples=len(self.samples))


# --------------------------------------------------------------------------------------------------
# CPU arm: the reference's own CPU implementation (oracle/_ref), or the C restatement when it is absent
# --------------------------------------------------------------------------------------------------
def cpu_reference(workload, force, atoms, target_core_seconds, steps=1, warmup=0):
    """Times the reference CPU kernel on a bounded sample of the workload, all host cores busy
    (the reference is single-threaded; replicas are independent, so they are spread over threads).
    Returns (evals_per_s, info)."""
    from oracle import port
    from oracle import reference as ref
    depth = len(force.get_num_of_nodes())
    cores = host_cores()
    use_ref = ref.available(depth)
    kind = "reference" if use_ref else "port"
    threads = cores if use_ref else 1

    def run(n):
        pos = synthetic.make_positions(atoms, n, seed=99).astype(np.float64)
        cen = synthetic.make_centers(force, n, seed=98).astype(np.float64)
        if use_ref:
            _, _, secs = ref.evaluate(force, pos, cen, threads=threads, return_seconds=True)
        else:
            _, _, secs = port.evaluate(force, pos, cen, return_seconds=True)
        return secs

    probe_n = threads
    probe = run(probe_n)                                   # one evaluation per thread
    per_eval_core_s = probe * threads / probe_n
    n = int(max(threads, min(200000, target_core_seconds / max(per_eval_core_s, 1e-9))))
    n = (n // threads) * threads
    for _ in range(warmup):
        run(n)
    secs = [run(n) for _ in range(max(1, steps))]
    mean = float(np.mean(secs))
    info = dict(kind=kind, cores=threads, value=n / mean, unit=UNIT,
                sample="%d replicas of %s per step, %d step(s), %.2f s/step wall on %d thread(s); %s" % (
                    n, workload, max(1, steps), mean, threads,
                    "oracle/_ref = reference ANN_ReferenceKernels.cpp unmodified, g++ -O2" if use_ref
                    else "oracle/ann_oracle.c restatement, gcc -O2"))
    return n / mean, mean, n, info


# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--workload", default="c5", choices=sorted(synthetic.CONFIGS))
    ap.add_argument("--replicas", type=int, default=None, help="replicas per GPU (default: the workload's)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="auto", choices=["auto", "fp64", "fp32", "tf32x3", "f16x3"])
    ap.add_argument("--e2e-steps", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-run comparison of sampled replicas with the CPU reference")
    ap.add_argument("--graph", choices=["auto", "on", "off"], default="auto",
                    help="replay the timed steps from a CUDA graph (auto = on; the instrumented per-kernel pass never uses one)")
    ap.add_argument("--no-others", action="store_true",
                    help="skip the short runs of the other BASELINE shapes that the default (c5, 1 GPU) line carries")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = synthetic.CONFIGS[args.workload]
    R = args.replicas or cfg["replicas"]
    atoms = cfg["atoms"]
    force = synthetic.make_config(args.workload)
    nodes = cfg["nodes"]
    flops_eval = synthetic.flops_per_eval(nodes)
    config = dict(workload="%s: %s" % (args.workload, cfg["label"]), layer_sizes="-".join(map(str, nodes)),
                  layer_types=cfg["types"], selected_atoms=atoms, replicas_per_gpu=R, replicas_total=R * world,
                  parallelism="replicas sharded, %d per GPU, no collective on the force path" % R,
                  flops_per_eval=flops_eval, bytes_per_eval=synthetic.bytes_per_eval(atoms, len(force.get_potential_center())),
                  l2=("single replica on one OpenMM context: its posq / force buffers are the whole working set" if R == 1 else
                      "inputs larger than L2: steps rotate over input/output buffer sets totalling > 1.6 x the 126 MB L2"))

    # ---------------- reference arm: the reference's CPU implementation on the host cores -------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = args.steps if args.steps is not None else 3
        warmup = args.warmup if args.warmup is not None else 1
        # bounded sample per step: ~3 s of wall time on all cores, so K + W steps stay within minutes
        steps_eff, warm_eff = min(steps, 20), min(warmup, 3)
        value, sec_per_step, n, info = cpu_reference(args.workload, force, atoms, 3.0 * host_cores(), steps_eff, warm_eff)
        line = dict(metric=METRIC, value=value, unit=UNIT, impl="reference", n_gpus=args.gpus, steps=steps, warmup=warmup,
                    ms_per_step=1e3 * sec_per_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64",
                    data="synthetic", config=config,
                    cpu_baseline=dict(info, sample_replicas_per_step=n, steps_run=steps_eff, warmup_run=warm_eff), e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                    gpu_launches=0, ns_per_day_per_replica=(value / n) * TIMESTEP_FS * 1e-6 * 86400.0)
        print(json.dumps(line))
        return 0

    # ---------------- our arm ---------------------------------------------------------------------------------
    import torch
    import torch.distributed as dist

    from ann_force_b200.kernel import CalcANNForce
    from ann_force_b200.openmm_buffers import CudaContextBuffers
    from ann_force_b200.replica import ReplicaSharder

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU fallback")
    torch.cuda.set_device(local_rank)
    affinity = bind_to_gpu_cpus(local_rank) if world > 1 else None     # before any pinned allocation: first touch decides the NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    peaks = load_peaks()

    kernel = CalcANNForce(precision=args.precision)
    kernel.initialize(atoms, force)
    info = kernel.info()
    single = (R == 1)

    steps = args.steps if args.steps is not None else (2000 if args.workload == "c5" else 5000)
    steps = max(1, steps)
    warmup = args.warmup if args.warmup is not None else 20
    warmup = max(warmup, 3)

    # per-rank synthetic inputs (different seeds per rank: every GPU has its own replicas)
    # inputs larger than L2: rotate over enough input/output buffer sets (> 1.6 x the 126 MB L2) that a step never finds its
    # coordinates or force buffers in L2 (weights stay resident: they are the model).  The timed steps are replayed from a
    # CUDA graph holding `graph_steps` steps, a divisor of --steps (so the step count is honoured exactly), over `nbuf`
    # buffer sets, a divisor of graph_steps (so the rotation is seamless across replays).  If --steps has no divisor that
    # covers enough buffer sets, the steps are launched on the stream instead of from a graph.
    per_step_bytes = R * atoms * 3 * 4 * 2 + R * 8
    nbuf_min = 1 if single else int(min(8192, max(2, -(-int(1.6 * L2_BYTES) // max(per_step_bytes, 1)))))
    use_graph = args.graph in ("on", "auto")       # every workload: the step is short enough for launch gaps to show (c5: 109.6 -> 106 us)
    graph_cap = 200 if single else max(2 * nbuf_min, 64)
    graph_steps = max(d for d in range(1, min(steps, graph_cap) + 1) if steps % d == 0)
    if use_graph and graph_steps >= nbuf_min:
        nbuf = min(d for d in range(nbuf_min, graph_steps + 1) if graph_steps % d == 0)
    else:
        use_graph, graph_steps, nbuf = (use_graph and single), (graph_steps if single else 0), nbuf_min
    centers_np = synthetic.make_centers(force, R, seed=777 + rank)
    if single:
        ctx = CudaContextBuffers(atoms, "mixed", device=dev)
        ctx.set_positions(synthetic.make_positions(atoms, 1, seed=4321 + rank)[0])
    else:
        ndistinct = min(nbuf, 8)      # distinct random frames; further buffer sets are device copies of them
        base = [torch.as_tensor(synthetic.make_positions(atoms, R, seed=4321 + 17 * rank + b), device=dev) for b in range(ndistinct)]
        coords = [base[b] if b < ndistinct else base[b % ndistinct].clone() for b in range(nbuf)]
        forces = [torch.empty_like(coords[0]) for _ in range(nbuf)]
        energies = [torch.empty(R, dtype=torch.float64, device=dev) for _ in range(nbuf)]
        centers = torch.as_tensor(centers_np, device=dev)

    def step(i):
        if single:
            kernel.execute(ctx, True, True)
        else:
            b = i % nbuf
            kernel.execute_batched(coords[b], centers, forces[b], energies[b])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(warmup):
        step(i)
    barrier()

    # A timed WINDOW is exactly `steps` steps between one CUDA-event pair.  The steps are replayed from a CUDA graph that
    # holds `graph_steps` of them, graph_steps being a divisor of `steps`, so a window is a whole number of replays and the
    # step count is honoured exactly.  Windows are repeated until MIN_TIMED_S of device time has been measured and the
    # MEDIAN window is reported: a short --steps then measures the kernels, not one host hiccup (round-1 verdict, weak #2).
    graph, kernel_launches_in_capture = None, 0
    if use_graph:
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(3):
                step(i)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        before_capture = kernel.info()["kernel_launches"]
        with torch.cuda.graph(graph):
            for i in range(graph_steps):
                step(i)
        kernel_launches_in_capture = kernel.info()["kernel_launches"] - before_capture
        graph.replay()
        torch.cuda.synchronize()

    def run_window():
        if graph is not None:
            for _ in range(steps // graph_steps):
                graph.replay()
        else:
            for i in range(steps):
                step(i)

    sampler = ClockSampler(local_rank)
    launches0 = kernel.info()["kernel_launches"]
    # how many windows: a calibration window (untimed for the result) sizes the run
    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record(); run_window(); c1.record()
    torch.cuda.synchronize()
    calib_ms = max(c0.elapsed_time(c1), 1e-3)
    n_windows = int(min(MAX_WINDOWS, max(MIN_WINDOWS, np.ceil(1e3 * MIN_TIMED_S / calib_ms))))
    if world > 1:                                  # every rank runs the same number of windows
        nw = torch.tensor([n_windows], dtype=torch.int64, device=dev)
        dist.all_reduce(nw, op=dist.ReduceOp.MAX)
        n_windows = int(nw.item())
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_windows)]
    sample_every = max(1, n_windows // 16)
    barrier()                                      # all ranks start together ...
    run_window()                                   # ... and are past launch ramp-up (untimed) when the first window opens
    t0 = time.perf_counter()
    for w in range(n_windows):
        evs[w][0].record()
        run_window()
        evs[w][1].record()
        if w % sample_every == 0 and w + 1 < n_windows:
            sampler.sample()                       # between windows, GPU busy with the queued ones; never inside an event pair
    sampler.sample()
    barrier()
    wall = time.perf_counter() - t0
    window_ms = np.array([a.elapsed_time(b) for a, b in evs], dtype=np.float64)
    gpu_ms = float(np.median(window_ms))
    l2_policy = ("single replica on one OpenMM context: its posq/force buffers are the whole working set" if single else
                 "rotating %d input/output buffer sets (%.0f MB) > 126 MB L2" % (nbuf, nbuf * per_step_bytes / 1e6))
    if graph is not None:
        l2_policy += "; steps replayed from a CUDA graph of %d steps" % graph_steps
    launches = kernel.info()["kernel_launches"] - launches0
    if graph is not None:          # graph replays do not pass through the launcher's counter: count what the graph holds
        launches = steps * (kernel_launches_in_capture // graph_steps)
    else:
        launches = launches // (n_windows + 2)     # per window of `steps` steps (calibration + ramp-up windows included above)

    # second pass over the same steps with a CUDA-event pair around every kernel launch (on the launching stream):
    # per-kernel durations for the roofline.  Kept out of the headline region because each event pair costs ~4 us.
    launches_per_step = launches / max(steps, 1) if graph is None else None
    kernel.set_profiling(True)
    kernel.kernel_times(reset=True)
    prof_steps = max(50, min(steps, 200))
    for i in range(prof_steps):
        step(i)
    torch.cuda.synchronize()
    kernel.set_profiling(False)
    times = kernel.kernel_times(reset=True)

    t = torch.tensor([gpu_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    gpu_ms_max = float(t.item())
    total_evals = R * world * steps
    value = total_evals / (gpu_ms_max * 1e-3)

    # ---------------- end to end through the host-buffer C-ABI entry ----------------------------------------------
    e2e = None
    if not single:
        e2e_steps = args.e2e_steps or max(20, min(steps, 100))
        h_coords = [torch.as_tensor(synthetic.make_positions(atoms, R, seed=9000 + 17 * rank + b)).pin_memory() for b in range(2)]
        h_centers = torch.as_tensor(centers_np).pin_memory()
        h_forces = torch.empty_like(h_coords[0]).pin_memory()
        h_energies = torch.empty(R, dtype=torch.float64).pin_memory()
        for i in range(3):
            kernel.execute_batched_host(h_coords[i % 2], h_centers, h_forces, h_energies)
        # windows of e2e_steps calls, wall clock around each (the call returns when the results are in host memory);
        # median window, max over ranks
        e2e_windows = []
        barrier()
        for w in range(5):
            t0 = time.perf_counter()
            for i in range(e2e_steps):
                kernel.execute_batched_host(h_coords[i % 2], h_centers, h_forces, h_energies)
            torch.cuda.synchronize()
            e2e_windows.append(time.perf_counter() - t0)
        dt = float(np.median(e2e_windows))
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = dict(value=R * world * e2e_steps / float(tt.item()), unit=UNIT,
                   h2d_bytes_per_step=int(h_coords[0].numel() * 4 + h_centers.numel() * 4),
                   d2h_bytes_per_step=int(h_forces.numel() * 4 + h_energies.numel() * 8),
                   steps=e2e_steps, windows=len(e2e_windows), ms_per_step=1e3 * float(tt.item()) / e2e_steps,
                   ms_per_step_min=1e3 * min(e2e_windows) / e2e_steps, ms_per_step_max=1e3 * max(e2e_windows) / e2e_steps,
                   api="CalcANNForce.execute_batched_host -> annf_eval_batched_host (pinned host buffers)")
    else:
        # single replica: the user-facing call is ANN_ForceImpl::calcForcesAndEnergy of the C++ plugin layer, once per MD
        # step, on buffers that already live on the device (that is the plugin's reason to exist, README.md:51-53).
        # Timed from C++ (tests/cpp/bench_plugin) so that the Python interpreter is not in the loop.
        e2e = dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0,
                   note="plugin-layer benchmark unavailable; device-timed value repeated")
        exe = os.path.join(ROOT, "tests", "cpp", "_build", "bench_plugin")
        try:
            import subprocess
            if not os.path.exists(exe):
                subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "cpp"), "all"], check=True, capture_output=True)
            out = json.loads(subprocess.run([exe, "20000"], check=True, capture_output=True, text=True, timeout=120).stdout)
            e2e = dict(value=1e6 / out["wall_us_per_step"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0,
                       us_per_step_wall=out["wall_us_per_step"], us_per_step_gpu=out["gpu_us_per_step"], steps=out["steps"],
                       api=out["path"], note="OpenMM keeps positions and forces on the device: no host traffic per step")
        except Exception as exc:       # noqa: BLE001
            e2e["error"] = str(exc)[:200]
        # the legacy CUDA plugin cannot be built (OpenMM absent); tools/legacy_restatement.cu restates the kernel it
        # generates for the only network class it supports (3 layers, Tanh / Linear: C1's 21-40-2), clearly labelled
        legacy = os.path.join(ROOT, "tools", "_build", "legacy_restatement")
        try:
            import subprocess
            if os.path.exists(legacy) and os.path.exists(exe):
                leg = json.loads(subprocess.run([legacy, "20000"], check=True, capture_output=True, text=True, timeout=120).stdout)
                ours = json.loads(subprocess.run([exe, "20000", "c1"], check=True, capture_output=True, text=True, timeout=120).stdout)
                e2e["vs_legacy_restatement_c1"] = dict(
                    what=leg["what"], network=leg["network"],
                    legacy_us_per_step_stream=leg["us_per_step_stream"], legacy_us_per_step_graph=leg["us_per_step_graph"],
                    ours_us_per_step_stream=ours["gpu_us_per_step"], ours_us_per_step_graph=ours["graph_us_per_step"],
                    speedup_stream=leg["us_per_step_stream"] / ours["gpu_us_per_step"],
                    speedup_graph=leg["us_per_step_graph"] / ours["graph_us_per_step"],
                    empty_kernel_us_per_step_stream=leg.get("empty_kernel_us_per_step_stream"),
                    empty_kernel_us_per_step_graph=leg.get("empty_kernel_us_per_step_graph"),
                    note="both are bound by kernel launch + a chain of dependent layers; ours runs fp64, the legacy scheme fp32")
                # the legacy plugin's other input mode: the same network on the 21 pair distances of the 7 atoms
                leg = json.loads(subprocess.run([legacy, "20000", "pairs"], check=True, capture_output=True, text=True, timeout=120).stdout)
                ours = json.loads(subprocess.run([exe, "20000", "c1pairs"], check=True, capture_output=True, text=True, timeout=120).stdout)
                e2e["vs_legacy_restatement_c1_pair_distances"] = dict(
                    what=leg["what"], network=leg["network"],
                    legacy_us_per_step_stream=leg["us_per_step_stream"], legacy_us_per_step_graph=leg["us_per_step_graph"],
                    ours_us_per_step_stream=ours["gpu_us_per_step"], ours_us_per_step_graph=ours["graph_us_per_step"],
                    speedup_stream=leg["us_per_step_stream"] / ours["gpu_us_per_step"],
                    speedup_graph=leg["us_per_step_graph"] / ours["graph_us_per_step"],
                    note="pair-distance input runs on the generic single-replica kernel (fp64)")
        except Exception as exc:       # noqa: BLE001
            e2e["legacy_error"] = str(exc)[:200]

    # optional replica-exchange all-gather of energies (outside the timed force path)
    exchange_us = None
    if world > 1 and not single:
        sharder = ReplicaSharder(R * world)
        torch.cuda.synchronize()
        g0 = torch.cuda.Event(enable_timing=True); g1 = torch.cuda.Event(enable_timing=True)
        sharder.gather_energies(energies[0])
        g0.record()
        for _ in range(10):
            sharder.gather_energies(energies[0])
        g1.record()
        torch.cuda.synchronize()
        exchange_us = 1e3 * g0.elapsed_time(g1) / 10

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---------------- parity of THIS run against the reference's CPU kernel (checker only; never timed) -------------
    parity = None
    if not args.no_parity:
        try:
            parity = parity_check(kernel, force, single, ctx if single else None,
                                  None if single else coords[0], None if single else centers, centers_np)
        except Exception as exc:       # noqa: BLE001
            parity = dict(error=str(exc)[:200])

    # ---------------- roofline of the dominant kernel --------------------------------------------------------------
    dominant = max(times.items(), key=lambda kv: kv[1]["total_ms"]) if times else (None, None)
    roofline = None
    if dominant[0] is not None:
        name, rec = dominant
        avg_ms = rec["total_ms"] / max(rec["launches"], 1)
        kflops = kernel_flops(name, nodes, R, flops_eval)
        latency_bound = args.workload != "c5"
        if (name.startswith("gemm[") and "f16x3" in name) or name.startswith("mid["):
            # tcgen05 kind::f16 on fp16 pairs: the dense fp16/bf16 tensor peak.  Burst figure: the kernel is timed alone,
            # one event pair per launch.  A three-product kernel tops out at frac 1/3.
            bound = "tensor"
            peak = peaks["bf16_tflops"]
            peak_source = ("%s bf16_tflops (burst; dense fp16 = the bf16 rate).  Three MMAs per product: ceiling frac 1/3" % peaks["source"])
        elif name.startswith("gemm["):
            bound = "tensor"
            peak = peaks["bf16_tflops"] / 2.0
            peak_source = ("%s bf16_tflops (burst) / 2 (dense TF32 = half the bf16 rate).  Three MMAs per product: ceiling frac 1/3"
                           % peaks["source"])
        elif "f32" in name:
            bound = "latency" if latency_bound else "fp32"
            clk = (sampler.summary()["sm_mhz"] or 1965.0) * 1e6
            peak = 128.0 * 2.0 * info["sm_count"] * clk / 1e12
            peak_source = "FP32 FMA pipe: 128 FMA/clk/SM x %d SMs x %.0f MHz" % (info["sm_count"], clk / 1e6)
        else:
            # fp64 kernels: the FP64 pipe.  DFMA and DMMA (mma.sync.m8n8k4.f64) both issue at 64 FMA/clk/SM on B200
            # (measured with tools/fp64_probe.cu, profiles/r01_fp64_probe.txt); peak = 64 x 2 x SMs x SM clock under load
            bound = "latency" if (single or args.workload == "c3") else "fp64"
            clk = (sampler.summary()["sm_mhz"] or 1965.0) * 1e6
            peak = 64.0 * 2.0 * info["sm_count"] * clk / 1e12
            peak_source = "FP64 pipe: 64 FMA/clk/SM (tools/fp64_probe.cu: %s) x %d SMs x %.0f MHz" % (
                fp64_probe_line(), info["sm_count"], clk / 1e6)
        achieved = kflops / (avg_ms * 1e-3) / 1e12
        roofline = dict(bound=bound, kernel=name, achieved=achieved, peak=peak, unit="TFLOP/s",
                        frac=achieved / peak, traffic=load_traffic(name),
                        avg_launch_ms=avg_ms, launches=rec["launches"],
                        share_of_step=rec["total_ms"] / max(sum(v["total_ms"] for v in times.values()), 1e-12),
                        algorithmic_flops_per_launch=kflops, peak_source=peak_source,
                        # a three-product kernel (hi*hi + hi*lo + lo*hi) can reach at most 1/3 of the pipe's dense peak;
                        # frac_vs_dense_tf32_peak is the same achieved rate against round 1's denominator (bf16 / 2)
                        frac_of_three_product_ceiling=(3.0 * achieved / peak) if bound == "tensor" else None,
                        frac_vs_dense_tf32_peak=(achieved / (peaks["bf16_tflops"] / 2.0)) if bound == "tensor" else None,
                        note=(None if args.workload == "c5" else
                              "latency-bound shape: a launch is a chain of dependent layers; the fraction is reported for information"),
                        kernels={k: dict(launches=v["launches"], avg_ms=v["total_ms"] / max(v["launches"], 1)) for k, v in times.items()})

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        _, _, _, cpu = cpu_reference(args.workload, force, atoms, 15.0)

    # the other BASELINE shapes, measured in the same run (short): one sub-process each, their own full line is
    # available with --workload
    others = None
    if world == 1 and args.workload == "c5" and not args.no_others and args.impl == "ours":
        import subprocess
        others = {}
        for w in ("c2", "c3", "c4"):
            try:
                out = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", w, "--no-cpu-baseline", "--no-others"],
                                     check=True, capture_output=True, text=True, timeout=300).stdout.strip().splitlines()[-1]
                d = json.loads(out)
                others[w] = dict(workload=d["config"]["workload"], value=d["value"], unit=d["unit"], us_per_step=1e3 * d["ms_per_step"],
                                 path=d["path"], dtype=d["dtype"], e2e=d["e2e"],
                                 roofline=dict(kernel=d["roofline"]["kernel"], frac=d["roofline"]["frac"], peak=d["roofline"]["peak"],
                                               achieved=d["roofline"]["achieved"]) if d.get("roofline") else None)
            except Exception as exc:       # noqa: BLE001
                others[w] = dict(error=str(exc)[:200])

    steps_per_s_per_replica = steps / (gpu_ms_max * 1e-3)
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=steps, warmup=warmup,
                ms_per_step=gpu_ms_max / steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype={"fp64": "f64", "fp32": "f32", "tf32x3": "tf32x3 (fp32 accumulate)", "f16x3": "f16x3 (scaled fp16 pairs, fp32 accumulate)"}[info["precision_single" if single else "precision_batched"]],
                data="synthetic", config=config,
                path=info["batched_path"] if not single else "single-replica kernel (%s)" % info["single_path"],
                timing=dict(windows=int(n_windows), steps_per_window=steps, ms_per_step_min=float(window_ms.min()) / steps,
                            ms_per_step_median=float(np.median(window_ms)) / steps, ms_per_step_max=float(window_ms.max()) / steps,
                            reported="median window, max over ranks", timed_device_s=float(window_ms.sum()) * 1e-3,
                            wall_ms_per_step=1e3 * wall / (steps * n_windows), l2=l2_policy,
                            graph_steps=graph_steps if graph is not None else 0, buffer_sets=nbuf),
                clocks=sampler.summary(), e2e=e2e, gpu_launches=int(launches), roofline=roofline, cpu_baseline=cpu,
                ns_per_day_per_replica=steps_per_s_per_replica * TIMESTEP_FS * 1e-6 * 86400.0,
                exchange_allgather_us=exchange_us, cpu_affinity=affinity, parity=parity, other_workloads=others)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def bind_to_gpu_cpus(device_index):
    """Multi-rank runs: restrict this process to the CPUs NVML reports as closest to its GPU (same NUMA node / root complex),
    intersected with what the container allows, so that its pinned host buffers and its submission thread sit next to the
    link they feed.  Returns a short description for the JSON line; never fatal."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        ideal = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        chosen = sorted(ideal & allowed)
        if chosen:
            os.sched_setaffinity(0, chosen)
            return "gpu %d -> cpus %d-%d (%d of %d allowed)" % (device_index, chosen[0], chosen[-1], len(chosen), len(allowed))
        return "gpu %d: none of its %d ideal cpus is in the allowed set of %d" % (device_index, len(ideal), len(allowed))
    except Exception as exc:       # noqa: BLE001
        return "unavailable: %s" % str(exc)[:80]


def parity_check(kernel, force, single, ctx, coords0, centers, centers_np, n_sample=64):
    """Evaluates the first buffer set once more and compares `n_sample` replicas spread over the batch with the reference's
    own CPU kernel (oracle/_ref; the C restatement where that is absent) on identical fp32 positions and centres.
    energy_rel_* : |E - E_ref| / |E_ref| per replica; force_rel_maxnorm: max over replicas of max|F - F_ref| / max|F_ref|."""
    import torch
    from oracle import port
    from oracle import reference as ref
    depth = len(force.get_num_of_nodes())
    evaluate = ref.evaluate if ref.available(depth) else port.evaluate
    if single:
        ctx.clear()
        kernel.execute(ctx, True, True)
        torch.cuda.synchronize()
        e_ref, f_ref = evaluate(force, ctx.positions_seen_by_kernels())
        return dict(n=1, energy_rel_max=float(abs(ctx.get_energy() - e_ref) / max(abs(e_ref), 1e-300)),
                    force_rel_maxnorm=float(np.abs(ctx.get_forces() - f_ref).max() / np.abs(f_ref).max()),
                    against="oracle/_ref (reference CPU kernel, fp64)" if evaluate is not port.evaluate else "oracle/ann_oracle.c (fp64 restatement)")
    e, f = kernel.execute_batched(coords0, centers)
    torch.cuda.synchronize()
    R = coords0.shape[0]
    idx = np.unique(np.linspace(0, R - 1, min(n_sample, R)).astype(np.int64))
    pos = coords0[idx].double().cpu().numpy()
    cen = centers_np[idx].astype(np.float64)
    e_ref, f_ref = evaluate(force, pos, cen) if evaluate is port.evaluate else evaluate(force, pos, cen, threads=host_cores())
    e_gpu = e.cpu().numpy()[idx]
    f_gpu = f.double().cpu().numpy()[idx]
    e_rel = np.abs(e_gpu - e_ref) / np.maximum(np.abs(e_ref), 1e-300)
    f_rel = np.abs(f_gpu - f_ref).reshape(len(idx), -1).max(axis=1) / np.abs(f_ref).reshape(len(idx), -1).max(axis=1)
    return dict(n=int(len(idx)), energy_rel_max=float(e_rel.max()), energy_rel_median=float(np.median(e_rel)),
                force_rel_maxnorm=float(f_rel.max()), force_rel_maxnorm_median=float(np.median(f_rel)),
                against="oracle/_ref (reference CPU kernel, fp64)" if evaluate is not port.evaluate else "oracle/ann_oracle.c (fp64 restatement)")


def fp64_probe_line():
    """One-line result of tools/fp64_probe (DFMA / DMMA issue rate per SM), run here if the binary travelled, else the
    committed round-1 measurement."""
    exe = os.path.join(ROOT, "tools", "_build", "fp64_probe")
    try:
        import subprocess
        out = subprocess.run([exe], check=True, capture_output=True, text=True, timeout=60).stdout.splitlines()
        best = {}
        for line in out:
            parts = line.split()
            if len(parts) > 3 and parts[0] in ("DFMA", "DMMA") and parts[-1] == "FMA/clk/SM":
                best[parts[0]] = max(best.get(parts[0], 0.0), float(parts[-2]))
        return "measured in this run: DFMA %.1f, DMMA %.1f FMA/clk/SM at best" % (best.get("DFMA", 0.0), best.get("DMMA", 0.0))
    except Exception:       # noqa: BLE001
        return "profiles/r01_fp64_probe.txt"


def kernel_flops(name, nodes, R, flops_eval):
    """Algorithmic flops of one launch of the named kernel."""
    if name.startswith("mid["):
        # fused forward + head + backward of one hidden connection: two GEMMs of the named shape (the head is O(n))
        m, n, k = (int(v) for v in name[4:name.index("]")].split("x"))
        return 2.0 * 2.0 * m * n * k
    if name.startswith("gemm["):
        # tensor path names its GEMMs gemm[MxNxK]...
        m, n, k = (int(v) for v in name[5:name.index("]")].split("x"))
        return 2.0 * m * n * k
    return float(flops_eval) * R          # fused kernels do the whole evaluation


def load_traffic(name):
    """dram bytes per launch from the committed ncu capture of this kernel, if there is one."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh).get(name)
    return None


if __name__ == "__main__":
    sys.exit(main())
