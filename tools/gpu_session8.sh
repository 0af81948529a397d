#!/bin/bash
# Round-2 GPU session 8: ncu --set full of every kernel of the C5 step (one step, after warm-up), bench, full GPU tests.
O=gpurun_out
mkdir -p $O
B="python bench.py --no-others --no-cpu-baseline"
echo "== bench default"; timeout -k 10 300 $B --steps 200 > $O/s8_bench_default.json 2> $O/s8_bench_default.err; tail -c 300 $O/s8_bench_default.err
echo "== pytest gpu (all)"; timeout -k 10 900 python -m pytest tests -q -m gpu --timeout 300 > $O/s8_pytest.log 2>&1; echo "rc=$?" >> $O/s8_pytest.log; tail -4 $O/s8_pytest.log
echo "== ncu full"
timeout -k 10 900 ncu --set full --clock-control none --import-source on --launch-skip 120 --launch-count 6 -o $O/r02_prof_c5_step -f \
   python bench.py --no-others --no-cpu-baseline --no-parity --steps 20 --warmup 3 --graph off --e2e-steps 1 > $O/s8_ncu.log 2>&1; tail -3 $O/s8_ncu.log
ls -la $O/r02_prof_c5_step.ncu-rep
echo done
