"""Per-phase SM-cycle breakdown of the single-replica kernel (diagnostic)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ann_force_b200 import synthetic, _capi
from ann_force_b200.kernel import CalcANNForce
from ann_force_b200.openmm_buffers import CudaContextBuffers
for cfg in ("c1", "c2", "c4"):
    force = synthetic.make_config(cfg)
    n = synthetic.CONFIGS[cfg]["atoms"]
    ctx = CudaContextBuffers(n, "mixed"); ctx.set_positions(synthetic.make_positions(n, 1)[0])
    k = CalcANNForce(); k.initialize(n, force)
    for _ in range(20): k.execute(ctx)
    buf = (C.c_int64 * 16)()
    _capi.check(_capi.lib().annf_debug_single_phases(k._handle, buf))
    full = np.array(buf[:11])
    print(cfg, "forward detail: matvec0", int(full[7] - full[2]), "act0", int(full[8] - full[7]), "matvec1", int(full[9] - full[8]), "act1", int(full[10] - full[9]))
    t = np.array(buf[:7]); d = np.diff(t)
    print(cfg, "cycles:", dict(zip(["features", "stage-wait", "forward", "head", "backward", "forces"], d.tolist())), "total", int(t[6] - t[0]))
