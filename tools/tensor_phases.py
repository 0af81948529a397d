"""Diagnostic: where the time of one C5 step goes on the tensor-core path, kernel by kernel and phase by phase.

Needs the phase-clock build of the library:
    make -C ann_force_b200/csrc EXTRA=-DANNF_PHASE_CLOCKS OUT=$PWD/ann_force_b200/libannforce_b200_clocks.so OBJDIR=$PWD/ann_force_b200/csrc/_obj_clocks
    ANNF_B200_LIB=$PWD/ann_force_b200/libannforce_b200_clocks.so python tools/tensor_phases.py [workload] [replicas]
%globaltimer marks (ns): per kernel, CTA 0's phase boundaries and the first / last CTA start and last CTA end of the grid.
Prints times relative to the moment CTA 0 of the step's first kernel entered.
"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from ann_force_b200 import _capi, synthetic
from ann_force_b200.kernel import CalcANNForce

work = sys.argv[1] if len(sys.argv) > 1 else "c5"
cfg = synthetic.CONFIGS[work]
R = int(sys.argv[2]) if len(sys.argv) > 2 else cfg["replicas"]
force = synthetic.make_config(work)
k = CalcANNForce(precision=os.environ.get("PHASE_PRECISION", "auto"))
k.initialize(cfg["atoms"], force)
coords = [torch.as_tensor(synthetic.make_positions(cfg["atoms"], R, seed=s), device="cuda") for s in range(8)]
cen = torch.as_tensor(synthetic.make_centers(force, R), device="cuda")
out = [torch.empty_like(coords[0]) for _ in range(8)]
en = torch.empty(R, dtype=torch.float64, device="cuda")
for i in range(16):
    k.execute_batched(coords[i % 8], cen, out[i % 8], en)
torch.cuda.synchronize()
names = ["k0", "k1", "k2", "k3", "k4", "k5", "k6", "k7"]   # launch order: prep, forward GEMMs, (fused middle | GEMM, head, GEMM), backward GEMMs
phase = ["entry", "prologue", "deps", "tma0", "landed0", "mma_done", "drained", "parked", "epilogue", "exit", "p10", "vectors", "staged", "stores_issued", "stores_read", "p15"]
buf = (C.c_int64 * 192)()
for rep in range(3):
    _capi.check(_capi.lib().annf_debug_tensor_phases(k._handle, buf, 1))
    # several steps back to back so that the recorded one runs under PDL overlap like the steady state; CTA-0 marks are
    # overwritten by every step, so what is read back is the LAST step (grid-wide min / max marks span all 4: not printed
    # for the start, printed for the end, which the last step owns)
    for i in range(4):
        k.execute_batched(coords[(rep + i) % 8], cen, out[(rep + i) % 8], en)
    _capi.check(_capi.lib().annf_debug_tensor_phases(k._handle, buf, 0))
    t = np.array(buf[:], dtype=np.int64).reshape(8, 24)
    t0 = min(int(t[s][0]) for s in range(8) if t[s][0] > 0)
    rows = {}
    print("--- repetition", rep)
    for s in range(8):
        if t[s][0] == 0:
            continue
        rec = {phase[i]: round((int(t[s][i]) - t0) / 1e3, 2) for i in range(16) if t[s][i] > 0}
        rec["last_cta_start"] = round((int(t[s][17]) - t0) / 1e3, 2)
        rec["last_cta_end"] = round((int(t[s][18]) - t0) / 1e3, 2)
        if t[s][19] or t[s][20] or t[s][21] or t[s][22]:       # SM cycles spent waiting, as us at 1.965 GHz (not times on the axis)
            print("      waits (us): mma on accumulators %.2f, mma on operands %.2f, producer on stages %.2f, epilogue on chunks %.2f" % tuple(int(t[s][i]) / 1965.0 for i in (19, 20, 21, 22)))
        rows[names[s]] = rec
        print("%-5s" % names[s], " ".join("%s=%.2f" % kv for kv in sorted(rec.items(), key=lambda kv: kv[1])))
print(json.dumps(dict(workload=work, replicas=R, unit="us since CTA 0 of the step's first kernel entered", kernels=rows)))
