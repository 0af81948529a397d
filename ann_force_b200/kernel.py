"""Host-side mirror of the reference's kernel object for the CUDA platform.

``CalcANNForce`` plays the role of ``CudaCalcANN_ForceKernel``
(platforms/cuda/include/CudaANN_Kernels.h:45-125): ``initialize`` / ``execute`` /
``copyParametersToContext`` with the argument meaning of ``CalcANN_ForceKernel``
(openmmapi/include/openmm/ANN_Kernels.h:36-53), plus the batched-replica entry the reference
does not have.  It is a thin layer over the C-ABI (include/annforce_b200.h); torch is used only
for device memory and streams.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _capi
from .force import ANN_Force


def _torch():
    import torch
    return torch


class CalcANNForce:
    @staticmethod
    def Name() -> str:                      # ANN_Kernels.h:23-25
        return "CalcANN_Force"

    def __init__(self, device: Optional[int] = None, precision: str = "auto"):
        self._handle = C.c_void_p()
        self._device = device
        self._precision = _capi.PRECISIONS[precision]
        self._num_center = 0
        self._num_system_atoms = 0

    # ---- CalcANN_ForceKernel::initialize(system, force) -----------------------------------------
    def initialize(self, num_system_atoms: int, force: ANN_Force) -> None:
        force.validate()
        torch = _torch()
        if not torch.cuda.is_available():
            raise _capi.ANNForceError(_capi.ERR_NO_DEVICE, "CalcANNForce.initialize: no CUDA device; there is no CPU fallback")
        device = torch.cuda.current_device() if self._device is None else int(self._device)
        L = force.num_layers()
        nodes = np.ascontiguousarray(force.get_num_of_nodes(), dtype=np.int32)
        types = np.ascontiguousarray([_capi.LAYER_CODES[t] for t in force.get_layer_types()], dtype=np.int32)
        coeff = [np.ascontiguousarray(w, dtype=np.float64).ravel() for w in force.get_coeffients_of_connections()]
        bias = [np.ascontiguousarray(b, dtype=np.float64).ravel() for b in force.get_values_of_biased_nodes()]
        center = np.ascontiguousarray(force.get_potential_center(), dtype=np.float64)
        atoms = np.ascontiguousarray(force.get_index_of_backbone_atoms(), dtype=np.int32)
        dihedrals = np.ascontiguousarray(force.get_list_of_index_of_atoms_forming_dihedrals(), dtype=np.int32).ravel()
        pairs = np.ascontiguousarray(force.get_list_of_pair_index_for_distances(), dtype=np.int32).ravel()
        as_d = lambda a: a.ctypes.data_as(_capi.c_double_p)
        as_i = lambda a: a.ctypes.data_as(_capi.c_int_p)
        coeff_p = (_capi.c_double_p * (L - 1))(*[as_d(w) for w in coeff])
        bias_p = (_capi.c_double_p * (L - 1))(*[as_d(b) for b in bias])
        desc = _capi.Desc(C.sizeof(_capi.Desc), device, L, as_i(nodes), as_i(types), coeff_p, bias_p,
                          as_d(center), int(center.size), force.get_force_constant(), force.get_scaling_factor(),
                          force.get_data_type_in_input_layer(), as_i(atoms), int(atoms.size),
                          as_i(dihedrals), int(dihedrals.size // 4), as_i(pairs), int(pairs.size // 2),
                          int(num_system_atoms), self._precision)
        self.close()
        _capi.check(_capi.lib().annf_create(C.byref(desc), C.byref(self._handle)))
        self._device = device
        self._num_center = int(center.size)
        self._num_system_atoms = int(num_system_atoms)
        # atoms whose force rows the kernels write: the Cartesian selection; for pair / dihedral inputs a subset we do not track here
        self._n_sel = int(atoms.size) if force.get_data_type_in_input_layer() == 1 else -1

    def close(self) -> None:
        if self._handle:
            _capi.lib().annf_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _need(self):
        if not self._handle:
            raise RuntimeError("CalcANNForce: initialize() has not been called")

    def info(self) -> dict:
        self._need()
        info = _capi.Info()
        _capi.check(_capi.lib().annf_get_info(self._handle, C.byref(info)))
        return dict(precision_single=_capi.PRECISION_NAMES[info.precision_single],
                    precision_batched=_capi.PRECISION_NAMES[info.precision_batched],
                    batched_path={0: "fused", 1: "tensor", 2: "stream", 3: "dmma"}[info.batched_path],
                    single_path="low-latency" if info.single_path == 1 else "generic",
                    cluster_size=info.cluster_size, macs_per_eval=info.macs_per_eval,
                    kernel_launches=info.kernel_launches, sm_count=info.sm_count)

    @staticmethod
    def _stream(stream):
        if stream is None:
            stream = _torch().cuda.current_stream()
        return C.c_void_p(int(getattr(stream, "cuda_stream", stream)))

    # ---- CalcANN_ForceKernel::execute(context, includeForces, includeEnergy) ---------------------------
    def execute(self, context, includeForces: bool = True, includeEnergy: bool = True, stream=None) -> float:
        """`context` is an OpenMM-CUDA-style buffer set (ann_force_b200.openmm_buffers.CudaContextBuffers).
        Like the reference's CUDA kernel (CudaANN_Kernels.cpp:367-369) this returns 0.0: the energy is
        accumulated into the context's device energy buffer and the forces into its fixed-point buffer."""
        self._need()
        flags = 0
        if context.posq.dtype == _torch().float64:
            flags |= _capi.FLAG_POSQ_DOUBLE
        if context.energy_buffer.dtype == _torch().float64:
            flags |= _capi.FLAG_ENERGY_DOUBLE
        if not includeForces:
            flags |= _capi.FLAG_SKIP_FORCES
        if not includeEnergy:
            flags |= _capi.FLAG_SKIP_ENERGY
        corr = context.posq_correction
        _capi.check(_capi.lib().annf_eval_openmm(
            self._handle, C.c_void_p(context.posq.data_ptr()), C.c_void_p(corr.data_ptr()) if corr is not None else None,
            C.c_void_p(context.force.data_ptr()), int(context.padded_num_atoms), C.c_void_p(context.energy_buffer.data_ptr()),
            flags, self._stream(stream)))
        return 0.0

    def set_atom_order(self, atom_index, stream=None) -> None:
        """atom_index: int32 device tensor, atom_index[slot] = original atom (CudaContext::getAtomIndexArray)."""
        self._need()
        _capi.check(_capi.lib().annf_set_atom_order(self._handle, C.c_void_p(atom_index.data_ptr()),
                                                    int(atom_index.numel()), self._stream(stream)))

    # ---- CalcANN_ForceKernel::copyParametersToContext(context, force) ----------------------------------------
    def copyParametersToContext(self, force: ANN_Force, stream=None) -> None:
        self._need()
        center = np.ascontiguousarray(force.get_potential_center(), dtype=np.float64)
        _capi.check(_capi.lib().annf_update_parameters(self._handle, center.ctypes.data_as(_capi.c_double_p), int(center.size),
                                                       float(force.get_force_constant()), self._stream(stream)))

    # ---- batched replicas (device tensors) -----------------------------------------------------------------------
    def execute_batched(self, coords, centers=None, forces=None, energies=None, stream=None):
        """coords: float32 CUDA tensor [R, N, 3]; centers: float32 [R, n_center] or None.
        Returns (energies float64 [R], forces float32 [R, N, 3]); rows of atoms outside the selection are
        left as they were in `forces` (zeros when allocated here)."""
        self._need()
        torch = _torch()
        assert coords.is_cuda and coords.dtype == torch.float32 and coords.is_contiguous() and coords.dim() == 3
        R, N, _ = coords.shape
        if forces is None:
            forces = torch.zeros_like(coords) if self._n_sel != N else torch.empty_like(coords)
        if energies is None:
            energies = torch.empty(R, dtype=torch.float64, device=coords.device)
        if centers is not None:
            assert centers.is_cuda and centers.dtype == torch.float32 and centers.is_contiguous()
            assert tuple(centers.shape) == (R, self._num_center)
        _capi.check(_capi.lib().annf_eval_batched(
            self._handle, int(R), C.c_void_p(coords.data_ptr()), int(N),
            C.c_void_p(centers.data_ptr()) if centers is not None else None,
            C.c_void_p(forces.data_ptr()), C.c_void_p(energies.data_ptr()), self._stream(stream)))
        return energies, forces

    # ---- batched replicas (HOST buffers: the end-to-end entry) --------------------------------------------------------
    def execute_batched_host(self, coords, centers=None, forces=None, energies=None):
        """coords: float32 [R, N, 3] numpy array or CPU torch tensor, C-contiguous (checked).  PINNED buffers let the copies
        overlap the evaluation of neighbouring chunks; pageable ones give the same result with the driver staging every copy.
        Copies in, evaluates, copies forces/energies out and waits.  Returns (energies, forces) of the same kind."""
        self._need()
        torch = _torch()
        is_torch = isinstance(coords, torch.Tensor)

        def ptr(x):
            return C.c_void_p(x.data_ptr() if is_torch else x.ctypes.data)

        def ok(x, dtype_t, dtype_n, shape, what):
            """The C entry takes raw pointers: a wrong dtype, a strided view or a device tensor would be misread silently."""
            if is_torch:
                good = isinstance(x, torch.Tensor) and not x.is_cuda and x.dtype == dtype_t and x.is_contiguous() and tuple(x.shape) == shape
            else:
                good = isinstance(x, np.ndarray) and x.dtype == dtype_n and x.flags["C_CONTIGUOUS"] and tuple(x.shape) == shape
            if not good:
                raise ValueError("execute_batched_host: %s must be a C-contiguous host %s of shape %s" % (what, dtype_n.__name__, shape))

        if coords.ndim != 3 or coords.shape[2] != 3:
            raise ValueError("execute_batched_host: coords must have shape [R, N, 3]")
        R, N, _ = coords.shape
        ok(coords, torch.float32, np.float32, (R, N, 3), "coords")
        if centers is not None:
            ok(centers, torch.float32, np.float32, (R, self._num_center), "centers")
        if forces is None:
            forces = torch.empty_like(coords) if is_torch else np.empty_like(coords)
        if energies is None:
            energies = torch.empty(R, dtype=torch.float64) if is_torch else np.empty(R, dtype=np.float64)
        ok(forces, torch.float32, np.float32, (R, N, 3), "forces")
        ok(energies, torch.float64, np.float64, (R,), "energies")
        _capi.check(_capi.lib().annf_eval_batched_host(self._handle, int(R), ptr(coords), int(N),
                                                       ptr(centers) if centers is not None else None,
                                                       ptr(forces), ptr(energies)))
        return energies, forces

    # ---- profiling ----------------------------------------------------------------------------------------------------
    def set_profiling(self, enabled: bool) -> None:
        self._need()
        _capi.check(_capi.lib().annf_set_profiling(self._handle, 1 if enabled else 0))

    def kernel_times(self, reset: bool = True) -> dict:
        self._need()
        buf = (_capi.KernelTime * 64)()
        n = C.c_int32(0)
        _capi.check(_capi.lib().annf_get_kernel_times(self._handle, buf, 64, C.byref(n), 1 if reset else 0))
        return {buf[i].name.decode(): dict(launches=int(buf[i].launches), total_ms=float(buf[i].total_ms))
                for i in range(min(n.value, 64))}
