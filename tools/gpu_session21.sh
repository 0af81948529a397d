#!/bin/bash
O=gpurun_out
mkdir -p $O
B="python bench.py --no-others --no-cpu-baseline"
echo "== bench default"; timeout -k 10 300 $B --steps 200 > $O/s21_bench_default.json 2> $O/s21_bench_default.err; tail -c 200 $O/s21_bench_default.err
echo "== bench no B prefetch"; ANNF_TENSOR_B_PREFETCH=0 timeout -k 10 300 $B --steps 200 --e2e-steps 5 --no-parity > $O/s21_bench_nobpre.json 2> $O/s21_bench_nobpre.err
echo done
