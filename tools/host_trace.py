"""Diagnostic: device timeline of annf_eval_batched_host at BASELINE C5 (ANNF_HOST_TRACE=1 makes the library print it).
    ANNF_HOST_TRACE=1 python tools/host_trace.py [chunks] [ratio]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ann_force_b200 import synthetic
from ann_force_b200.kernel import CalcANNForce
cfg = synthetic.CONFIGS["c5"]
force = synthetic.make_config("c5")
k = CalcANNForce(); k.initialize(cfg["atoms"], force)
R = cfg["replicas"]
pos = torch.as_tensor(synthetic.make_positions(cfg["atoms"], R)).pin_memory()
cen = torch.as_tensor(synthetic.make_centers(force, R)).pin_memory()
f = torch.empty_like(pos).pin_memory(); e = torch.empty(R, dtype=torch.float64).pin_memory()
for i in range(6):
    k.execute_batched_host(pos, cen, f, e)
