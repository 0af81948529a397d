#!/bin/bash
# Round-2 GPU session 27: prep flavours test; warp prep with coalesced stores.
O=gpurun_out
mkdir -p $O
echo "== pytest prep flavours"; timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -k "prep_kernel_flavours" > $O/s27_pytest.log 2>&1; echo "rc=$?" >> $O/s27_pytest.log; tail -4 $O/s27_pytest.log
for mode in 0 1 0 1; do
echo "== bench k20 prep mode $mode"; ANNF_PREP_MODE=$mode timeout -k 10 600 python bench.py --steps 20 --warmup 5 2>/dev/null | cut -c1-200
done
export ANNF_B200_LIB=$PWD/ann_force_b200/libannforce_b200_clocks.so
echo "== phases mode 1"; ANNF_PREP_MODE=1 timeout -k 10 300 python tools/tensor_phases.py c5 > $O/s27_phases_mode1.txt 2>&1; grep "^k0" $O/s27_phases_mode1.txt | head -3
echo "== phases mode 0"; timeout -k 10 300 python tools/tensor_phases.py c5 > $O/s27_phases_mode0.txt 2>&1; grep "^k0" $O/s27_phases_mode0.txt | head -3
echo done
