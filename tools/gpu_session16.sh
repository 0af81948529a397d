#!/bin/bash
O=gpurun_out
mkdir -p $O
B="python bench.py --no-others --no-cpu-baseline"
echo "== bench default"; timeout -k 10 300 $B --steps 200 > $O/s16_bench_default.json 2> $O/s16_bench_default.err; tail -c 200 $O/s16_bench_default.err
echo "== bench prep128"; ANNF_PREP_THREADS=128 timeout -k 10 300 $B --steps 200 --e2e-steps 5 --no-parity > $O/s16_bench_prep128.json 2> $O/s16_bench_prep128.err
echo "== pytest tensor"; timeout -k 10 600 python -m pytest tests -q -m gpu --timeout 120 -x -k "tensor or wide_network or f16" > $O/s16_pytest.log 2>&1; echo "rc=$?" >> $O/s16_pytest.log; tail -3 $O/s16_pytest.log
echo done
