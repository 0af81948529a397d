#!/bin/bash
# Round-2 GPU session 4: fused middle kernel (fwd + head + bwd of the last hidden connection).
O=gpurun_out
mkdir -p $O
echo "== pytest mid"; timeout -k 10 600 python -m pytest tests -q -m gpu --timeout 300 -x -k "tensor_core_path_matches or fused_middle or wide_network or f16" > $O/s4_pytest_mid.log 2>&1; echo "rc=$?" >> $O/s4_pytest_mid.log; tail -15 $O/s4_pytest_mid.log
echo "== pytest gpu (all)"; timeout -k 10 900 python -m pytest tests -q -m gpu --timeout 300 > $O/s4_pytest.log 2>&1; echo "rc=$?" >> $O/s4_pytest.log; tail -4 $O/s4_pytest.log
B="python bench.py --no-others --no-cpu-baseline"
echo "== bench f16 (mid)"; timeout -k 10 300 $B --steps 200 > $O/s4_bench_f16.json 2> $O/s4_bench_f16.err
echo "== bench f16 (no mid)"; ANNF_TENSOR_MID=0 timeout -k 10 300 $B --steps 200 --e2e-steps 5 > $O/s4_bench_f16_nomid.json 2> $O/s4_bench_f16_nomid.err
echo "== bench f16 pair tiles S=3"; ANNF_GEMM_BN=130 timeout -k 10 300 $B --steps 200 --e2e-steps 5 > $O/s4_bench_f16_bn130.json 2> $O/s4_bench_f16_bn130.err
export ANNF_B200_LIB=$PWD/ann_force_b200/libannforce_b200_clocks.so
echo "== phases f16"; timeout -k 10 300 python tools/tensor_phases.py c5 > $O/s4_phases_f16.txt 2>&1; tail -7 $O/s4_phases_f16.txt | head -6
echo done
