#!/bin/bash
# Round-2 GPU session 9: prep at 8 CTAs / SM (one wave); full GPU tests; default bench incl. other workloads and CPU baseline.
O=gpurun_out
mkdir -p $O
echo "== pytest gpu (all)"; timeout -k 10 900 python -m pytest tests -q -m gpu --timeout 300 > $O/s9_pytest.log 2>&1; echo "rc=$?" >> $O/s9_pytest.log; tail -4 $O/s9_pytest.log
echo "== bench k20"; timeout -k 10 600 python bench.py --steps 20 --warmup 5 > $O/s9_bench_k20.json 2> $O/s9_bench_k20.err; tail -c 300 $O/s9_bench_k20.err
export ANNF_B200_LIB=$PWD/ann_force_b200/libannforce_b200_clocks.so
echo "== phases"; timeout -k 10 300 python tools/tensor_phases.py c5 > $O/s9_phases_f16.txt 2>&1; tail -8 $O/s9_phases_f16.txt | head -7
echo done
