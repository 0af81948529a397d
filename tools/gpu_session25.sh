#!/bin/bash
O=gpurun_out
mkdir -p $O
B="python bench.py --no-others --no-cpu-baseline"
echo "== pytest tensor"; timeout -k 10 600 python -m pytest tests -q -m gpu --timeout 120 -x -k "tensor or wide_network or f16 or split_k" > $O/s25_pytest.log 2>&1; echo "rc=$?" >> $O/s25_pytest.log; tail -3 $O/s25_pytest.log
echo "== bench default"; timeout -k 10 300 $B --steps 200 > $O/s25_bench_default.json 2> $O/s25_bench_default.err; tail -c 200 $O/s25_bench_default.err
echo "== bench tf32"; timeout -k 10 300 $B --steps 200 --precision tf32x3 --e2e-steps 5 > $O/s25_bench_tf32.json 2> $O/s25_bench_tf32.err
export ANNF_B200_LIB=$PWD/ann_force_b200/libannforce_b200_clocks.so
echo "== phases"; timeout -k 10 300 python tools/tensor_phases.py c5 > $O/s25_phases.txt 2>&1; tail -13 $O/s25_phases.txt | grep "^k3"
echo done
