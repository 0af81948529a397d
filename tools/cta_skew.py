#!/usr/bin/env python
"""Per-CTA timeline of the force GEMM (the last kernel of a C5 step): when does every CTA finish its main loop, its TMEM
drain, its TMA stores, and when does it exit?  Phase-clock build only (ANNF_B200_LIB=..._clocks.so).

    ANNF_B200_LIB=$PWD/ann_force_b200/libannforce_b200_clocks.so python tools/cta_skew.py
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ann_force_b200 import _capi, synthetic  # noqa: E402
from ann_force_b200.kernel import CalcANNForce  # noqa: E402

slot = int(sys.argv[1]) if len(sys.argv) > 1 else 5          # kernel of the step: 1 fwd0, 2 fwd1, 4 bwd1, 5 bwd0 (force GEMM)
cfg = synthetic.CONFIGS["c5"]
R = cfg["replicas"] if "replicas" in cfg else 1024
force = synthetic.make_config("c5")
k = CalcANNForce(precision="auto")
k.initialize(cfg["atoms"], force)
coords = [torch.as_tensor(synthetic.make_positions(cfg["atoms"], R, seed=s), device="cuda") for s in range(8)]
cen = torch.as_tensor(synthetic.make_centers(force, R), device="cuda")
out = [torch.empty_like(coords[0]) for _ in range(8)]
en = torch.empty(R, dtype=torch.float64, device="cuda")
for i in range(8):
    k.execute_batched(coords[i % 8], cen, out[i % 8], en)
torch.cuda.synchronize()
assert _capi.lib().annf_debug_cta_clocks(None, slot) == 0
for i in range(24):
    k.execute_batched(coords[i % 8], cen, out[i % 8], en)
torch.cuda.synchronize()
buf = (C.c_longlong * (10 * 1024 + 8 * 256))()
rc = _capi.lib().annf_debug_cta_clocks(buf, slot)
assert rc == 0, rc
raw = np.array(buf[:], dtype=np.int64)
t = raw[:10 * 1024].reshape(10, 1024)
sm_exit = raw[10 * 1024:].reshape(8, 256)
n = int((t[3] > 0).sum())
smid = t[5, :n]
t = t[:, :n].astype(np.float64)
t0 = t[3].min() - 40e3
prev_exit = sm_exit[{5: 4, 4: 2, 2: 1, 1: 5}[slot]]      # the GEMM that held the SMs before this one
cols = [(6, "entry"), (7, "prologue"), (8, "deps"), (4, "operands"), (0, "main done"), (1, "drained"), (2, "stores read"), (3, "exit")]
gx = 36 if slot == 5 else 4
print("kernel slot", slot, "CTAs:", n, " (us on an arbitrary common axis; operands / main done: the MMA issuer lives in the even CTA of a pair)")
for i, nm in cols:
    v = (t[i][t[i] > t0] - t0) / 1e3
    if v.size == 0:
        continue
    print("%-12s min %6.2f  p10 %6.2f  median %6.2f  p90 %6.2f  max %6.2f   (n=%d)" % (nm, v.min(), np.percentile(v, 10), np.median(v), np.percentile(v, 90), v.max(), v.size))
print("per CTA: block (x, y) sm | previous GEMM left this SM | " + ", ".join(nm for _, nm in cols))
for b in range(n):
    pe = prev_exit[smid[b]]
    print("%3d (%2d,%d) sm %3d | %6s | %s" % (b, b % gx, b // gx, smid[b], ("%6.2f" % ((pe - t0) / 1e3)) if pe > t0 else "  -  ",
                                          " ".join("%6.2f" % ((t[i][b] - t0) / 1e3) if t[i][b] > t0 else "   -  " for i, _ in cols)))
