#!/bin/bash
O=gpurun_out
mkdir -p $O
export ANNF_B200_LIB=$PWD/ann_force_b200/libannforce_b200_clocks.so
echo "== phases default"; timeout -k 10 300 python tools/tensor_phases.py c5 > $O/s14_phases.txt 2>&1; tail -14 $O/s14_phases.txt | head -13
echo "== phases bn130"; ANNF_GEMM_BN=130 timeout -k 10 300 python tools/tensor_phases.py c5 > $O/s14_phases_bn130.txt 2>&1; tail -14 $O/s14_phases_bn130.txt | head -5
echo "== phases chunk2"; ANNF_TENSOR_CHUNK=2 timeout -k 10 300 python tools/tensor_phases.py c5 > $O/s14_phases_chunk2.txt 2>&1; tail -14 $O/s14_phases_chunk2.txt | head -13
echo done
