#!/bin/bash
# Round-2 GPU session 5: where the fused middle kernel spends its time; prep with 128 / 256 threads; host-entry chunk counts.
O=gpurun_out
mkdir -p $O
export ANNF_B200_LIB=$PWD/ann_force_b200/libannforce_b200_clocks.so
echo "== phases f16"; timeout -k 10 300 python tools/tensor_phases.py c5 > $O/s5_phases_f16.txt 2>&1; tail -6 $O/s5_phases_f16.txt | head -5
echo "== phases f16 prep256"; ANNF_PREP_THREADS=256 timeout -k 10 300 python tools/tensor_phases.py c5 > $O/s5_phases_f16_prep256.txt 2>&1; tail -6 $O/s5_phases_f16_prep256.txt | head -2
unset ANNF_B200_LIB
B="python bench.py --no-others --no-cpu-baseline --no-parity"
for c in 2 4 6 8; do
  echo "== host chunks $c"; ANNF_TENSOR_MID=0 ANNF_HOST_CHUNKS=$c timeout -k 10 300 $B --steps 20 --e2e-steps 100 > $O/s5_bench_chunks$c.json 2> $O/s5_bench_chunks$c.err
done
echo done
