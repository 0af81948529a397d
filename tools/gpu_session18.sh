#!/bin/bash
O=gpurun_out
mkdir -p $O
export ANNF_B200_LIB=$PWD/ann_force_b200/libannforce_b200_clocks.so
timeout -k 10 200 python tools/single_phases.py 2>&1 | tee $O/s18_single_phases.txt
echo done
