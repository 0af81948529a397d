"""Small end-to-end exercise of every kernel path, meant to be run under compute-sanitizer where that is allowed
(memcheck / racecheck / synccheck):   compute-sanitizer --tool memcheck python tools/sanitize_smoke.py [single|dmma|stream|tensor]
(The build pool of this repository keeps compute-sanitizer closed, so there it only serves as a quick smoke run.)
Sizes are tiny on purpose; parity is asserted loosely (the -m gpu tests hold the real bars)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ann_force_b200 import synthetic
from ann_force_b200.kernel import CalcANNForce
from ann_force_b200.openmm_buffers import CudaContextBuffers
from oracle import port

def single(cfg):
    force = synthetic.make_config(cfg); n = synthetic.CONFIGS[cfg]["atoms"]
    ctx = CudaContextBuffers(n, "mixed"); ctx.set_positions(synthetic.make_positions(n, 1)[0])
    k = CalcANNForce(); k.initialize(n, force); k.execute(ctx); torch.cuda.synchronize()
    e_ref, f_ref = port.evaluate(force, ctx.positions_seen_by_kernels())
    assert abs(ctx.get_energy() - e_ref) <= 1e-9 * abs(e_ref), cfg
    print("single", cfg, k.info()["single_path"], "ok")

def batched(nodes, types, atoms, R, precision):
    force = synthetic.make_force(nodes, types, atoms, seed=5)
    pos = synthetic.make_positions(atoms, R, seed=6); cen = synthetic.make_centers(force, R, seed=7)
    k = CalcANNForce(precision=precision); k.initialize(atoms, force)
    e, f = k.execute_batched(torch.as_tensor(pos, device="cuda"), torch.as_tensor(cen, device="cuda")); torch.cuda.synchronize()
    e_ref, _ = port.evaluate(force, pos.astype(np.float64), cen.astype(np.float64))
    np.testing.assert_allclose(e.cpu().numpy(), e_ref, rtol=1e-3)
    print("batched", nodes, precision, k.info()["batched_path"], "ok")

which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "single"):
    single("c2"); single("c4")
if which in ("all", "dmma"):
    batched([21, 40, 4], ["Tanh", "Circular"], 7, 19, "fp64")
    batched([180, 100, 100, 3], ["Tanh", "Tanh", "Tanh"], 60, 21, "fp64")
if which in ("all", "stream"):
    batched([180, 100, 100, 3], ["Tanh", "Tanh", "Tanh"], 60, 21, "fp32")
if which in ("all", "tensor"):
    batched([1200, 320, 256, 6], ["Tanh", "Tanh", "Tanh"], 400, 130, "tf32x3")
