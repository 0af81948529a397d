#!/bin/bash
O=gpurun_out
mkdir -p $O
echo "== pytest gpu"; timeout -k 10 1200 python -m pytest tests -q -m gpu --timeout 300 > $O/s17_pytest.log 2>&1; echo "rc=$?" >> $O/s17_pytest.log; tail -6 $O/s17_pytest.log
echo "== bench c2"; timeout -k 10 300 python bench.py --workload c2 --no-cpu-baseline > $O/s17_bench_c2.json 2> $O/s17_bench_c2.err
echo "== cpp plugin test"; timeout -k 10 200 tests/cpp/_build/test_ann_plugin > $O/s17_cpp.log 2>&1; tail -2 $O/s17_cpp.log
echo done
