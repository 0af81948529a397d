#!/bin/bash
# Round-2 GPU session 13: why are the 256 x 128 CTA-pair tiles (ANNF_GEMM_BN=130) slow on the K = 4500 GEMM?  phases + ncu.  Also C4 in fp32.
O=gpurun_out
mkdir -p $O
echo "== c4 fp32"; timeout -k 10 300 python bench.py --workload c4 --precision fp32 --no-cpu-baseline --no-others > $O/s13_bench_c4_fp32.json 2> $O/s13_bench_c4_fp32.err
echo "== c4 fp64"; timeout -k 10 300 python bench.py --workload c4 --no-cpu-baseline --no-others > $O/s13_bench_c4_fp64.json 2> $O/s13_bench_c4_fp64.err
echo "== c3 fp32"; timeout -k 10 300 python bench.py --workload c3 --precision fp32 --no-cpu-baseline --no-others > $O/s13_bench_c3_fp32.json 2> $O/s13_bench_c3_fp32.err
export ANNF_B200_LIB=$PWD/ann_force_b200/libannforce_b200_clocks.so
echo "== phases bn130"; ANNF_GEMM_BN=130 timeout -k 10 300 python tools/tensor_phases.py c5 > $O/s13_phases_bn130.txt 2>&1; tail -8 $O/s13_phases_bn130.txt | head -3
unset ANNF_B200_LIB
echo "== ncu bn130 fwd0"
ANNF_GEMM_BN=130 timeout -k 10 600 ncu --set full --clock-control none --import-source on --launch-skip 121 --launch-count 1 -o $O/r02_prof_c5_fwd0_bn130 -f \
   python bench.py --no-others --no-cpu-baseline --no-parity --steps 20 --warmup 3 --graph off --e2e-steps 1 > $O/s13_ncu.log 2>&1; tail -2 $O/s13_ncu.log
echo done
