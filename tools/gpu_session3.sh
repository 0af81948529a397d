#!/bin/bash
# Round-2 GPU session 3: epilogue rework (staged vectors, rolled loop, TMA-store force tiles), new prep, chunk = 1.
O=gpurun_out
mkdir -p $O
echo "== pytest gpu (tensor tests first)"; timeout 900 python -m pytest tests -q -m gpu --timeout 600 -x -k "tensor or wide or f16" > $O/s3_pytest_tensor.log 2>&1; echo "rc=$?" >> $O/s3_pytest_tensor.log; tail -4 $O/s3_pytest_tensor.log
echo "== pytest gpu (all)"; timeout 900 python -m pytest tests -q -m gpu --timeout 600 > $O/s3_pytest.log 2>&1; echo "rc=$?" >> $O/s3_pytest.log; tail -4 $O/s3_pytest.log
B="python bench.py --no-others --no-cpu-baseline"
echo "== bench f16"; timeout 300 $B --steps 200 > $O/s3_bench_f16.json 2> $O/s3_bench_f16.err
echo "== bench f16 no direct store"; ANNF_TENSOR_DIRECT_STORE=0 timeout 300 $B --steps 200 --e2e-steps 5 > $O/s3_bench_f16_nodirect.json 2> $O/s3_bench_f16_nodirect.err
echo "== bench tf32"; ANNF_AUTO_TENSOR=tf32x3 timeout 300 $B --steps 200 --e2e-steps 5 > $O/s3_bench_tf32.json 2> $O/s3_bench_tf32.err
export ANNF_B200_LIB=$PWD/ann_force_b200/libannforce_b200_clocks.so
echo "== phases f16"; timeout 300 python tools/tensor_phases.py c5 > $O/s3_phases_f16.txt 2>&1; tail -8 $O/s3_phases_f16.txt
echo done
