"""(Needs a diagnostic build of the library: make -C ann_force_b200/csrc clean && make -C ann_force_b200/csrc EXTRA=-DANNF_PHASE_CLOCKS)
Per-phase SM-cycle breakdown of the streamed batched kernel, CTA 0 (diagnostic)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ann_force_b200 import synthetic, _capi
from ann_force_b200.kernel import CalcANNForce
for cfg in ("c3", "c4"):
    force = synthetic.make_config(cfg)
    c = synthetic.CONFIGS[cfg]
    n, R = c["atoms"], c["replicas"]
    k = CalcANNForce(); k.initialize(n, force)
    coords = torch.as_tensor(synthetic.make_positions(n, R), device="cuda")
    cen = torch.as_tensor(synthetic.make_centers(force, R), device="cuda")
    for _ in range(20): k.execute_batched(coords, cen)
    buf = (C.c_int64 * 64)()
    _capi.check(_capi.lib().annf_debug_stream_phases(k._handle, buf))
    L = len(c["nodes"]); npass = 2 * (L - 1)
    if len(sys.argv) > 1:
        import time
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(200): k.execute_batched(coords, cen)
        torch.cuda.synchronize(); print(cfg, "us per call (stream of 200):", 1e6 * (time.perf_counter() - t0) / 200)
    t = np.array(buf[:5 + npass])
    names = ["prologue", "features"] + ["fwd%d" % l for l in range(L - 1)] + ["head"] + ["bwd%d" % l for l in range(L - 2, -1, -1)] + ["forces"]
    order = [0, 1, 2] + [3 + p for p in range(L - 1)] + [3 + npass] + [3 + p for p in range(L - 1, npass)] + [4 + npass]
    ts = t[order]
    print(cfg, k.info()["batched_path"], dict(zip(names, np.diff(ts).tolist())), "total", int(ts[-1] - ts[0]))
