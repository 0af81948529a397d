#!/bin/bash
# Round-2 GPU session 28: bulk-copy prep kernel with 1, 2, 4, 8 warps per replica.
O=gpurun_out
mkdir -p $O
echo "== pytest prep flavours"; timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -k "prep_kernel_flavours" > $O/s28_pytest.log 2>&1; echo "rc=$?" >> $O/s28_pytest.log; tail -4 $O/s28_pytest.log
for w in 1 2 4 8 1 2 4 8; do
echo "== bench k20 prep warps $w"; ANNF_PREP_WARPS=$w timeout -k 10 600 python bench.py --steps 20 --warmup 5 2>/dev/null | cut -c1-200
done
export ANNF_B200_LIB=$PWD/ann_force_b200/libannforce_b200_clocks.so
for w in 1 2 4 8; do
echo "== phases warps $w"; ANNF_PREP_WARPS=$w timeout -k 10 300 python tools/tensor_phases.py c5 > $O/s28_phases_w$w.txt 2>&1; grep "^k0" $O/s28_phases_w$w.txt | head -3
done
echo done
