"""B200-native ANN bias force (drop-in for the ANN_Force OpenMM plugin's hot path).

Public surface:
  ANN_Force        host-side parameter object, same setters as the reference / its SWIG wrapper
  CalcANNForce     the kernel object: initialize / execute / copyParametersToContext over the C-ABI
  ReplicaSharder   shards independent replicas across ranks (one process per GPU)
"""
from .force import ANN_Force, INPUT_CARTESIAN, INPUT_COS_SIN_OF_DIHEDRALS, INPUT_PAIR_DISTANCES, LAYER_TYPES

__all__ = ["ANN_Force", "INPUT_CARTESIAN", "INPUT_COS_SIN_OF_DIHEDRALS", "INPUT_PAIR_DISTANCES", "LAYER_TYPES"]
