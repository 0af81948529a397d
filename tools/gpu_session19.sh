#!/bin/bash
O=gpurun_out
mkdir -p $O
echo "== pytest gpu"; timeout -k 10 1200 python -m pytest tests -q -m gpu --timeout 300 > $O/s19_pytest.log 2>&1; echo "rc=$?" >> $O/s19_pytest.log; tail -4 $O/s19_pytest.log
echo "== cpp plugin test"; timeout -k 10 200 tests/cpp/_build/test_ann_plugin > $O/s19_cpp.log 2>&1; tail -2 $O/s19_cpp.log
echo "== bench c2"; timeout -k 10 300 python bench.py --workload c2 --no-cpu-baseline > $O/s19_bench_c2.json 2> $O/s19_bench_c2.err
echo "== bench c3"; timeout -k 10 300 python bench.py --workload c3 --no-cpu-baseline > $O/s19_bench_c3.json 2> $O/s19_bench_c3.err
ANNF_B200_LIB=$PWD/ann_force_b200/libannforce_b200_clocks.so timeout -k 10 200 python tools/single_phases.py 2>&1 | tee $O/s19_single_phases.txt | grep "c2"
echo done
