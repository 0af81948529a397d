#!/bin/bash
# Round-2 GPU session 6: pipelined prep, one-wave head, split-K 4 for the short-K GEMMs.
O=gpurun_out
mkdir -p $O
echo "== pytest tensor"; timeout -k 10 600 python -m pytest tests -q -m gpu --timeout 300 -x -k "tensor or fused_middle or wide_network or f16" > $O/s6_pytest_tensor.log 2>&1; echo "rc=$?" >> $O/s6_pytest_tensor.log; tail -5 $O/s6_pytest_tensor.log
B="python bench.py --no-others --no-cpu-baseline"
echo "== bench default"; timeout -k 10 300 $B --steps 200 > $O/s6_bench_default.json 2> $O/s6_bench_default.err
echo "== bench old prep"; ANNF_PREP_PIPELINED=0 timeout -k 10 300 $B --steps 200 --e2e-steps 5 --no-parity > $O/s6_bench_oldprep.json 2> $O/s6_bench_oldprep.err
echo "== bench old prep 256"; ANNF_PREP_PIPELINED=0 ANNF_PREP_THREADS=256 timeout -k 10 300 $B --steps 200 --e2e-steps 5 --no-parity > $O/s6_bench_oldprep256.json 2> $O/s6_bench_oldprep256.err
echo "== bench shortk 4"; ANNF_GEMM_SPLIT_SHORTK=4 timeout -k 10 300 $B --steps 200 --e2e-steps 5 --no-parity > $O/s6_bench_shortk4.json 2> $O/s6_bench_shortk4.err
echo "== bench shortk 1"; ANNF_GEMM_SPLIT_SHORTK=1 timeout -k 10 300 $B --steps 200 --e2e-steps 5 --no-parity > $O/s6_bench_shortk1.json 2> $O/s6_bench_shortk1.err
export ANNF_B200_LIB=$PWD/ann_force_b200/libannforce_b200_clocks.so
echo "== phases"; timeout -k 10 300 python tools/tensor_phases.py c5 > $O/s6_phases_f16.txt 2>&1; tail -8 $O/s6_phases_f16.txt | head -7
echo "== phases shortk4"; ANNF_GEMM_SPLIT_SHORTK=4 timeout -k 10 300 python tools/tensor_phases.py c5 > $O/s6_phases_f16_shortk4.txt 2>&1; tail -8 $O/s6_phases_f16_shortk4.txt | head -7
echo done
