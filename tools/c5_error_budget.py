#!/usr/bin/env python
"""Where does the energy error of the tensor-core path come from?  (round-1 verdict, weak #1)

CPU-only numerical experiment (numpy): BASELINE C5 (4500-512-512-8, Tanh) is evaluated in fp64 — the reference's arithmetic —
and then again with ONE source of rounding switched on at a time, in the order the device path applies them:

  split      operands kept as an 11 + 11-bit pair (TF32 pair, or fp16 pair of the row-scaled value), lo*lo dropped
  chunk      fp32 accumulation: every --chunk-k wide K chunk summed exactly, rounded to fp32 (the TMEM accumulator's best case),
             then added into an fp32 accumulator with round-to-nearest (the "promotion" adds of the kernel)
  trunc      the same, but the in-chunk accumulation truncates toward zero after every MMA instruction (an emulation of the
             non-rounding tensor-core accumulator; 8 (tf32) or 16 (f16) products per instruction, three instructions per step)
  act        tanh and 1 - a^2 evaluated in fp32 (tanh itself taken as correctly rounded: tanhf is within 2 ulp of that)
  all        everything together = a model of the device path

Output: per setting, the median and maximum relative energy error and the force max-norm error over the sampled replicas.
This is an explanation of the measured numbers in bench.py's `parity` block, not a substitute for them.

    python tools/c5_error_budget.py [--replicas 48] [--kind f16|tf32]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ann_force_b200 import synthetic  # noqa: E402


def rn_bits(v, bits):
    """round-to-nearest to `bits` significant bits (fp64 in, fp64 out)"""
    m, e = np.frexp(v)
    return np.ldexp(np.round(m * (1 << bits)), e - bits)


def split_tf32(v):
    v = v.astype(np.float32).astype(np.float64)
    hi = rn_bits(v, 11)
    lo = rn_bits(v - hi, 11)
    return hi, lo, None


def row_exp(bound):
    e = np.zeros(bound.shape, dtype=np.int64)
    ok = bound > 0
    q = np.frexp(bound[ok])[1]
    e[ok] = np.clip(14 - q, -60, 60)
    return e


def split_f16(v, exps):
    """v [rows, cols] fp64, exps [rows] -> (hi, lo) of v * 2^e as fp16 values held in fp64, and the exponents"""
    s = np.ldexp(1.0, exps)[:, None]
    vs = (v.astype(np.float32).astype(np.float64) * s).astype(np.float32)
    hi = vs.astype(np.float16)
    lo = (vs - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float64), lo.astype(np.float64), exps


def trunc32(x):
    """fp64 -> fp32 truncated toward zero"""
    f = x.astype(np.float32)
    over = np.abs(f.astype(np.float64)) > np.abs(x)
    f[over] = np.nextafter(f[over], np.float32(0))
    return f


def contract(a, w, kind, mode, a_exp=None, w_exp=None, chunk_k=64):
    """a [R, K] · w[N, K]^T the way the device does it.  mode: exact | split | chunk | trunc"""
    if mode == "exact":
        return a @ w.T
    if kind == "tf32":
        ah, al, _ = split_tf32(a)
        wh, wl, _ = split_tf32(w)
        unscale = 1.0
        kstep = 8
    else:
        ah, al, _ = split_f16(a, a_exp)
        wh, wl, _ = split_f16(w, w_exp)
        unscale = np.ldexp(1.0, -(a_exp[:, None] + w_exp[None, :]))
        kstep = 16
    K = a.shape[1]
    if mode == "split":
        return (al @ wh.T + ah @ wl.T + ah @ wh.T) * unscale
    acc = np.zeros((a.shape[0], w.shape[0]), dtype=np.float32)
    for k0 in range(0, K, chunk_k):
        sl = slice(k0, min(K, k0 + chunk_k))
        if mode == "chunk":
            part = (al[:, sl] @ wh[:, sl].T + ah[:, sl] @ wl[:, sl].T + ah[:, sl] @ wh[:, sl].T).astype(np.float32)
        else:
            part = np.zeros_like(acc)
            for k1 in range(sl.start, sl.stop, kstep):
                s2 = slice(k1, min(sl.stop, k1 + kstep))
                for x, y in ((al, wh), (ah, wl), (ah, wh)):
                    part = trunc32(part.astype(np.float64) + x[:, s2] @ y[:, s2].T)
        acc = (acc + part).astype(np.float32)          # IEEE round-to-nearest add, as the promotion does
    return acc.astype(np.float64) * unscale


def evaluate(force, pos, cen, kind, mode, act32, chunk_k=64):
    nodes = force.get_num_of_nodes()
    L = len(nodes)
    W = [np.asarray(w, dtype=np.float64).reshape(nodes[l + 1], nodes[l]) for l, w in enumerate(force.get_coeffients_of_connections())]
    B = [np.asarray(b, dtype=np.float64) for b in force.get_values_of_biased_nodes()]
    k, s = force.get_force_constant(), force.get_scaling_factor()
    R, A = pos.shape[0], pos.shape[1]
    x = pos.astype(np.float64)
    x = ((x - x.mean(axis=1, keepdims=True)) / s).reshape(R, 3 * A)
    if mode != "exact":
        x = x.astype(np.float32).astype(np.float64)
    f32 = (lambda v: v.astype(np.float32).astype(np.float64)) if act32 else (lambda v: v)
    acts = [x]
    bound = np.abs(x).max(axis=1)
    for l in range(L - 2):
        a_exp = row_exp(bound)
        w_exp = row_exp(np.abs(W[l].astype(np.float32)).max(axis=1).astype(np.float64))
        z = contract(acts[-1], W[l], kind, mode, a_exp, w_exp, chunk_k)
        z = f32(f32(z) + f32(B[l]))
        acts.append(f32(np.tanh(z)))
        bound = np.ones(R)
    a_last = acts[-1]
    if mode != "exact":
        # the head reads the 22-bit pair
        h, lo, _ = split_tf32(a_last) if kind == "tf32" else split_f16(a_last, row_exp(np.ones(R)))
        a_last = (h + lo) if kind == "tf32" else (h + lo) * np.ldexp(1.0, -row_exp(np.ones(R)))[:, None]
    cv = np.tanh(a_last @ W[L - 2].T + B[L - 2])
    e = 0.5 * k * ((cv - cen) ** 2).sum(axis=1)
    dz = k * (cv - cen) * (1 - cv * cv)
    g = dz.astype(np.float32).astype(np.float64) @ W[L - 2]
    if act32:
        g = f32(f32(g) * f32(1 - f32(a_last * a_last)))
    else:
        g = g * (1 - a_last * a_last)
    for l in range(L - 3, -1, -1):
        Wt = W[l].T.copy()                              # [n_l][n_{l+1}]
        if l == 0:
            wsum = np.stack([W[0][:, c::3].sum(axis=1) for c in range(3)], axis=1)      # [n_1][3]
            Wc = W[0] - np.tile(wsum, (1, A)) / A
            Wt = (-(Wc) / s).T.copy()
        gmax = np.abs(g).max(axis=1)
        a_exp = row_exp(gmax * 1.0)
        w_exp = row_exp(np.abs(Wt.astype(np.float32)).max(axis=1).astype(np.float64))
        g = contract(g, Wt, kind, mode, a_exp, w_exp, chunk_k)
        if l > 0:
            a = acts[l]
            g = f32(f32(g) * f32(1 - f32(a * a))) if act32 else g * (1 - a * a)
    return e, g.reshape(R, A, 3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--replicas", type=int, default=32)
    ap.add_argument("--kind", default="f16", choices=["f16", "tf32"])
    ap.add_argument("--workload", default="c5")
    ap.add_argument("--chunk-k", type=int, default=64, help="K values accumulated in the (truncating) tensor-core accumulator between promotions")
    args = ap.parse_args()
    cfg = synthetic.CONFIGS[args.workload]
    force = synthetic.make_config(args.workload)
    pos = synthetic.make_positions(cfg["atoms"], args.replicas, seed=4321)
    cen = synthetic.make_centers(force, args.replicas, seed=777).astype(np.float64)
    e0, f0 = evaluate(force, pos, cen, args.kind, "exact", False)
    out = {}
    for name, mode, act32 in (("split", "split", False), ("chunk", "chunk", False), ("trunc", "trunc", False),
                              ("act", "exact", True), ("all", "trunc", True)):
        e, f = evaluate(force, pos, cen, args.kind, mode, act32, args.chunk_k)
        rel = np.abs(e - e0) / np.abs(e0)
        frel = np.abs(f - f0).reshape(args.replicas, -1).max(axis=1) / np.abs(f0).reshape(args.replicas, -1).max(axis=1)
        out[name] = dict(energy_rel_median=float(np.median(rel)), energy_rel_max=float(rel.max()), force_rel_maxnorm=float(frel.max()))
        print("%-6s energy rel median %.2e max %.2e   force max-norm %.2e" % (name, np.median(rel), rel.max(), frel.max()), flush=True)
    print(json.dumps(dict(workload=args.workload, kind=args.kind, replicas=args.replicas, chunk_k=args.chunk_k, settings=out)))


if __name__ == "__main__":
    main()
