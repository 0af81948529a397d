"""Device buffers laid out the way OpenMM's CUDA platform lays them out, for driving the
single-replica entry (annf_eval_openmm) without OpenMM (which is absent from this image).

Conventions reproduced [OpenMM API knowledge, SURVEY.md §7 "OpenMM CUDA buffer conventions"]:
  * posq: float4 per atom slot (single / mixed precision) or double4 (double); mixed mode carries the
    low-order bits in a second float4 array (posqCorrection);
  * atoms may be re-ordered: atom_index[slot] = original atom index;
  * forces: signed 64-bit fixed point, scale 2**32, laid out [x: padded][y: padded][z: padded];
  * energy: an accumulation buffer of float (single) or double (mixed / double); kernels add into it.
"""
from __future__ import annotations

import numpy as np
import torch

FORCE_SCALE = float(1 << 32)


class CudaContextBuffers:
    def __init__(self, num_atoms: int, precision: str = "single", device=None, order=None, energy_slots: int = 64):
        assert precision in ("single", "mixed", "double")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.num_atoms = int(num_atoms)
        self.padded_num_atoms = ((self.num_atoms + 31) // 32) * 32      # OpenMM pads to its tile size
        self.precision = precision
        pos_dtype = torch.float64 if precision == "double" else torch.float32
        self.posq = torch.zeros(self.padded_num_atoms, 4, dtype=pos_dtype, device=self.device)
        self.posq_correction = (torch.zeros(self.padded_num_atoms, 4, dtype=torch.float32, device=self.device)
                                if precision == "mixed" else None)
        self.force = torch.zeros(3 * self.padded_num_atoms, dtype=torch.int64, device=self.device)
        self.energy_buffer = torch.zeros(energy_slots, dtype=torch.float32 if precision == "single" else torch.float64,
                                         device=self.device)
        order = np.arange(self.num_atoms) if order is None else np.asarray(order)
        assert sorted(order.tolist()) == list(range(self.num_atoms))
        self._order = order.astype(np.int64)                              # slot -> original atom
        self.atom_index = torch.as_tensor(order.astype(np.int32), device=self.device)

    def set_positions(self, positions) -> None:
        """positions: [num_atoms, 3] in ORIGINAL atom order (nm), float64 or float32."""
        pos = np.asarray(positions, dtype=np.float64)[self._order]       # slot order
        full = np.zeros((self.padded_num_atoms, 4), dtype=np.float64)
        full[: self.num_atoms, :3] = pos
        if self.precision == "double":
            self.posq.copy_(torch.as_tensor(full))
        else:
            hi = full.astype(np.float32)
            self.posq.copy_(torch.as_tensor(hi))
            if self.posq_correction is not None:
                self.posq_correction.copy_(torch.as_tensor((full - hi.astype(np.float64)).astype(np.float32)))

    def positions_seen_by_kernels(self) -> np.ndarray:
        """[num_atoms, 3] float64 in original order: exactly the values the device buffers hold."""
        p = self.posq.double().cpu().numpy()[: self.num_atoms, :3]
        if self.posq_correction is not None:
            p = p + self.posq_correction.double().cpu().numpy()[: self.num_atoms, :3]
        out = np.zeros_like(p)
        out[self._order] = p
        return out

    def clear(self) -> None:
        self.force.zero_()
        self.energy_buffer.zero_()

    def get_forces(self) -> np.ndarray:
        """[num_atoms, 3] float64 in original order (kJ/mol/nm), decoded from fixed point."""
        raw = self.force.cpu().numpy().reshape(3, self.padded_num_atoms)[:, : self.num_atoms].T
        out = np.zeros((self.num_atoms, 3))
        out[self._order] = raw.astype(np.float64) / FORCE_SCALE
        return out

    def get_energy(self) -> float:
        return float(self.energy_buffer.double().sum().item())
