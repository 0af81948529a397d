"""Helpers for the -m gpu parity tests: run the CUDA path through the public kernel object."""
import numpy as np
import torch

from ann_force_b200.kernel import CalcANNForce
from ann_force_b200.openmm_buffers import CudaContextBuffers


def run_single(force, positions, precision="single", order=None, kernel_precision="auto", repeats=1):
    """One evaluation on OpenMM-layout buffers.  Returns (energy, forces[N,3], positions the kernel saw)."""
    n = positions.shape[0]
    ctx = CudaContextBuffers(n, precision=precision, order=order)
    ctx.set_positions(positions)
    kernel = CalcANNForce(precision=kernel_precision)
    kernel.initialize(n, force)
    if order is not None:
        kernel.set_atom_order(ctx.atom_index)
    for _ in range(repeats):
        assert kernel.execute(ctx, True, True) == 0.0
    torch.cuda.synchronize()
    return ctx.get_energy(), ctx.get_forces(), ctx.positions_seen_by_kernels(), kernel


def run_batched(force, positions, centers=None, precision="auto", host=False):
    """positions [R,N,3] (any float dtype; rounded to fp32 like the device input).  Returns
    (energies f64[R], forces f64[R,N,3], kernel)."""
    pos32 = np.ascontiguousarray(positions, dtype=np.float32)
    cen32 = None if centers is None else np.ascontiguousarray(centers, dtype=np.float32)
    kernel = CalcANNForce(precision=precision)
    kernel.initialize(pos32.shape[1], force)
    if host:
        f = np.full_like(pos32, np.nan)
        e, f = kernel.execute_batched_host(pos32, cen32, forces=f)
        return np.asarray(e, dtype=np.float64), np.asarray(f, dtype=np.float64), kernel
    coords = torch.as_tensor(pos32, device="cuda")
    cen = None if cen32 is None else torch.as_tensor(cen32, device="cuda")
    e, f = kernel.execute_batched(coords, cen)
    torch.cuda.synchronize()
    return e.cpu().numpy(), f.double().cpu().numpy(), kernel


def max_norm_rel(got, want):
    """Relative max-norm error per replica (the north-star force criterion), worst replica."""
    got = np.asarray(got, dtype=np.float64).reshape(want.shape[0] if want.ndim == 3 else 1, -1)
    want = np.asarray(want, dtype=np.float64).reshape(got.shape)
    return float(np.max(np.abs(got - want).max(axis=1) / np.maximum(np.abs(want).max(axis=1), 1e-300)))
