#!/usr/bin/env python
"""Would two half-batches on two streams beat one batch on one stream?  (diagnostic, BASELINE C5)

The C5 step is a chain of six kernels; three of them (last hidden connection forward, head, the same connection backward)
are latency-bound single-wave kernels that leave most of the machine idle.  If two independent half-size steps interleave on
two streams, the idle parts of one could be filled by the other.  This probe measures the aggregate rate of

  one     one handle, R replicas, one stream                         (what bench.py measures)
  two     two handles, R/2 replicas each, two streams, same graph    (fork/join inside the captured graph)

both replayed from CUDA graphs of 24 steps over rotating buffer sets larger than L2.  Prints one JSON line.

    python tools/two_stream_probe.py [--replicas 1024]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ann_force_b200 import synthetic  # noqa: E402
from ann_force_b200.kernel import CalcANNForce  # noqa: E402


def build(force, atoms, R, nbuf, seed):
    k = CalcANNForce(precision="auto")
    k.initialize(atoms, force)
    coords = [torch.as_tensor(synthetic.make_positions(atoms, R, seed=seed + b), device="cuda") for b in range(nbuf)]
    forces = [torch.empty_like(coords[0]) for _ in range(nbuf)]
    energies = [torch.empty(R, dtype=torch.float64, device="cuda") for _ in range(nbuf)]
    centers = torch.as_tensor(synthetic.make_centers(force, R, seed=777 + seed), device="cuda")
    return k, coords, forces, energies, centers


def time_graph(capture, steps_in_graph, replays=60):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        capture(3)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        capture(steps_in_graph)
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    times = []
    for _ in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(replays):
            g.replay()
        b.record()
        torch.cuda.synchronize()
        times.append(a.elapsed_time(b) * 1e3 / (replays * steps_in_graph))
    return float(np.median(times))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--replicas", type=int, default=1024)
    args = ap.parse_args()
    cfg = synthetic.CONFIGS["c5"]
    atoms, R = cfg["atoms"], args.replicas
    force = synthetic.make_config("c5")
    nbuf, steps = 8, 24
    one = build(force, atoms, R, nbuf, 100)
    halves = [build(force, atoms, R // 2, nbuf, 200), build(force, atoms, R // 2, nbuf, 300)]
    second = torch.cuda.Stream()

    def run_one(n):
        k, c, f, e, cen = one
        for i in range(n):
            k.execute_batched(c[i % nbuf], cen, f[i % nbuf], e[i % nbuf])

    def run_two(n):
        cur = torch.cuda.current_stream()
        second.wait_stream(cur)
        for i in range(n):
            for (k, c, f, e, cen), s in zip(halves, (cur, second)):
                k.execute_batched(c[i % nbuf], cen, f[i % nbuf], e[i % nbuf], stream=s)
        cur.wait_stream(second)

    us_one = time_graph(run_one, steps)
    us_two = time_graph(run_two, steps)
    print(json.dumps(dict(workload="c5", replicas=R, us_per_step_one_stream=us_one, us_per_step_two_half_batches=us_two,
                          evals_per_s_one=R / us_one * 1e6, evals_per_s_two=R / us_two * 1e6)))


if __name__ == "__main__":
    main()
