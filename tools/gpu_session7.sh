#!/bin/bash
# Round-2 GPU session 7: fp32 reference-atom prep, one-wave head, split-K 4 for the short-K GEMMs (fused middle kernel off).
O=gpurun_out
mkdir -p $O
echo "== pytest tensor"; timeout -k 10 600 python -m pytest tests -q -m gpu --timeout 300 -x -k "tensor or fused_middle or wide_network or f16" > $O/s7_pytest_tensor.log 2>&1; echo "rc=$?" >> $O/s7_pytest_tensor.log; tail -5 $O/s7_pytest_tensor.log
B="python bench.py --no-others --no-cpu-baseline"
echo "== bench default"; timeout -k 10 300 $B --steps 200 > $O/s7_bench_default.json 2> $O/s7_bench_default.err
echo "== bench prep 128"; ANNF_PREP_THREADS=128 timeout -k 10 300 $B --steps 200 --e2e-steps 5 --no-parity > $O/s7_bench_prep128.json 2> $O/s7_bench_prep128.err
echo "== bench prep pipelined"; ANNF_PREP_PIPELINED=1 timeout -k 10 300 $B --steps 200 --e2e-steps 5 --no-parity > $O/s7_bench_preppipe.json 2> $O/s7_bench_preppipe.err
echo "== bench shortk 4"; ANNF_GEMM_SPLIT_SHORTK=4 timeout -k 10 300 $B --steps 200 --e2e-steps 5 --no-parity > $O/s7_bench_shortk4.json 2> $O/s7_bench_shortk4.err
echo "== bench shortk 1"; ANNF_GEMM_SPLIT_SHORTK=1 timeout -k 10 300 $B --steps 200 --e2e-steps 5 --no-parity > $O/s7_bench_shortk1.json 2> $O/s7_bench_shortk1.err
export ANNF_B200_LIB=$PWD/ann_force_b200/libannforce_b200_clocks.so
echo "== phases"; timeout -k 10 300 python tools/tensor_phases.py c5 > $O/s7_phases_f16.txt 2>&1; tail -8 $O/s7_phases_f16.txt | head -7
echo done
