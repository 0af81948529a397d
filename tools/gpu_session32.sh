#!/bin/bash
# Round-2 GPU session 32: compile-time L2 eviction hints, one library per mask (ann_force_b200/_variants/, built with
# make EXTRA=-DANNF_L2_HINTS=n OUT=...); the default library is mask 15.
O=gpurun_out
mkdir -p $O
run() { timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-others --no-cpu-baseline --no-parity 2>/dev/null | python -c 'import json,sys; print(json.load(sys.stdin)["ms_per_step"])'; }
for rep in 1 2; do
echo "mask 15 (default lib): $(run)"
for h in 0 1 5 8 13 7; do
  echo "mask $h: $(ANNF_B200_LIB=$PWD/ann_force_b200/_variants/libannforce_b200_h$h.so run)"
done
done > $O/s32_hints.txt 2>&1; cat $O/s32_hints.txt
timeout -k 10 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
   --launch-skip 240 --launch-count 12 --csv --log-file $O/s32_dram_steady_h15.csv \
   python bench.py --no-others --no-cpu-baseline --no-parity --steps 60 --warmup 3 --graph off --e2e-steps 1 > $O/s32_ncu.log 2>&1
echo "== tests"; timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -k "tensor or f16 or tf32 or prep" 2>&1 | tail -2
echo done
