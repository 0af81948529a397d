#!/bin/bash
# Round-2 GPU session 26: coalesced prep walk (default) vs 4-atom groups (ANNF_PREP_MODE=2) vs warp per replica (1).
O=gpurun_out
mkdir -p $O
echo "== pytest tensor"; timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -k "tensor or f16 or tf32 or wide or split or c5" > $O/s26_pytest.log 2>&1; echo "rc=$?" >> $O/s26_pytest.log; tail -4 $O/s26_pytest.log
echo "== bench k20 warp prep"; timeout -k 10 600 python bench.py --steps 20 --warmup 5 > $O/s26_bench_k20.json 2> $O/s26_bench_k20.err; tail -c 300 $O/s26_bench_k20.err; cut -c1-200 $O/s26_bench_k20.json
echo "== bench k20 cta prep"; ANNF_PREP_MODE=2 timeout -k 10 600 python bench.py --steps 20 --warmup 5 > $O/s26_bench_k20_mode1.json 2> $O/s26_bench_k20_mode1.err; cut -c1-200 $O/s26_bench_k20_mode1.json
echo "== bench k20 warp"; ANNF_PREP_MODE=1 timeout -k 10 600 python bench.py --steps 20 --warmup 5 > $O/s26_bench_k20_mode2.json 2>/dev/null; cut -c1-200 $O/s26_bench_k20_mode2.json
echo "== bench k20 tf32"; timeout -k 10 600 python bench.py --steps 20 --warmup 5 --precision tf32x3 > $O/s26_bench_k20_tf32.json 2> $O/s26_bench_k20_tf32.err; cut -c1-200 $O/s26_bench_k20_tf32.json
export ANNF_B200_LIB=$PWD/ann_force_b200/libannforce_b200_clocks.so
echo "== phases"; timeout -k 10 300 python tools/tensor_phases.py c5 > $O/s26_phases_f16.txt 2>&1; tail -8 $O/s26_phases_f16.txt | head -7
echo done
