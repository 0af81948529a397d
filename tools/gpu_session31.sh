#!/bin/bash
# Round-2 GPU session 31: L2 eviction-priority hints (ANNF_L2_HINTS bit mask: 1 coords in evict_first, 2 forces out evict_first,
# 4 input operand pair evict_last, 8 weights evict_last).
O=gpurun_out
mkdir -p $O
for h in 0 1 2 4 8 5 9 12 13 15 0 15; do
  echo "hints $h: $(ANNF_L2_HINTS=$h timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-others --no-cpu-baseline --no-parity 2>/dev/null | python -c 'import json,sys; print(json.load(sys.stdin)["ms_per_step"])')"
done > $O/s31_hints.txt 2>&1; cat $O/s31_hints.txt
for h in 15; do
ANNF_L2_HINTS=$h timeout -k 10 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
   --launch-skip 240 --launch-count 12 --csv --log-file $O/s31_dram_steady_h$h.csv \
   python bench.py --no-others --no-cpu-baseline --no-parity --steps 60 --warmup 3 --graph off --e2e-steps 1 > $O/s31_ncu.log 2>&1
done
echo "== tests"; ANNF_L2_HINTS=15 timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -q -m gpu --timeout 300 -k "tensor or f16 or tf32 or prep" 2>&1 | tail -2
echo done
