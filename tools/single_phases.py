"""(Needs a diagnostic build of the library: make -C ann_force_b200/csrc clean && make -C ann_force_b200/csrc EXTRA=-DANNF_PHASE_CLOCKS)
Per-phase SM-cycle breakdown of the single-replica kernel (diagnostic)."""
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
    print(cfg, k.info()["single_path"], "forward detail: connection 0 (+activation)", int(full[7] - full[2]), "connection 1", int(full[9] - full[7]))
    full = np.array(buf[:16])
    print(cfg, "prologue: copies issued", int(full[11] - full[0]), "features done", int(full[1] - full[11]))
    t = np.array(buf[:7]); d = np.diff(t)
    print(cfg, "cycles:", dict(zip(["features", "stage-wait", "forward", "head", "backward", "forces"], d.tolist())), "total", int(t[6] - t[0]))
