#!/bin/bash
# Round-2 GPU session 2: TMEM chunk-length experiment (accuracy), phase clocks of the C5 step.
O=gpurun_out
mkdir -p $O
B="python bench.py --no-others --no-cpu-baseline"
for c in 1 2 4 8; do
  echo "== chunk $c f16"; ANNF_TENSOR_CHUNK=$c timeout 300 $B --steps 100 --e2e-steps 5 > $O/s2_bench_f16_chunk$c.json 2> $O/s2_bench_f16_chunk$c.err
  echo "== chunk $c tf32"; ANNF_AUTO_TENSOR=tf32x3 ANNF_TENSOR_CHUNK=$c timeout 300 $B --steps 100 --e2e-steps 5 > $O/s2_bench_tf32_chunk$c.json 2> $O/s2_bench_tf32_chunk$c.err
done
export ANNF_B200_LIB=$PWD/ann_force_b200/libannforce_b200_clocks.so
echo "== phases f16"; timeout 300 python tools/tensor_phases.py c5 > $O/s2_phases_f16.txt 2>&1; tail -8 $O/s2_phases_f16.txt
echo "== phases tf32"; ANNF_AUTO_TENSOR=tf32x3 timeout 300 python tools/tensor_phases.py c5 > $O/s2_phases_tf32.txt 2>&1
echo "== phases f16 R=128"; timeout 300 python tools/tensor_phases.py c5 128 > $O/s2_phases_f16_r128.txt 2>&1
echo done
