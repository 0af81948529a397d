#!/bin/bash
# Round-2 GPU session 15: split-K through global memory (CTA-pair tiles for the K = 4500 GEMM), chunk 2 for 256-wide tiles.
O=gpurun_out
mkdir -p $O
echo "== pytest gsplit"; timeout -k 10 600 python -m pytest tests -q -m gpu --timeout 120 -x -k "split_k_through or tensor_core_path_matches or wide_network" > $O/s15_pytest_gs.log 2>&1; echo "rc=$?" >> $O/s15_pytest_gs.log; tail -5 $O/s15_pytest_gs.log
B="python bench.py --no-others --no-cpu-baseline"
echo "== bench default"; timeout -k 10 300 $B --steps 200 > $O/s15_bench_default.json 2> $O/s15_bench_default.err; tail -c 200 $O/s15_bench_default.err
echo "== bench no gsplit"; ANNF_GEMM_GSPLIT=0 timeout -k 10 300 $B --steps 200 --e2e-steps 5 > $O/s15_bench_nogs.json 2> $O/s15_bench_nogs.err
for g in 6 8; do echo "== bench gsplit $g"; ANNF_GEMM_GSPLIT=$g timeout -k 10 300 $B --steps 200 --e2e-steps 5 --no-parity > $O/s15_bench_gs$g.json 2> $O/s15_bench_gs$g.err; done
export ANNF_B200_LIB=$PWD/ann_force_b200/libannforce_b200_clocks.so
echo "== phases"; timeout -k 10 300 python tools/tensor_phases.py c5 > $O/s15_phases.txt 2>&1; tail -13 $O/s15_phases.txt | head -12
echo done
