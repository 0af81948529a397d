#!/bin/bash
# Round-2 GPU session 30: steady-state DRAM traffic per kernel (no cache flush between kernels, single-pass metrics).
O=gpurun_out
mkdir -p $O
timeout -k 10 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
   --launch-skip 240 --launch-count 36 --csv --log-file $O/s30_dram_steady.csv \
   python bench.py --no-others --no-cpu-baseline --no-parity --steps 60 --warmup 3 --graph off --e2e-steps 1 > $O/s30_ncu.log 2>&1
tail -3 $O/s30_ncu.log | cut -c1-200
grep -c . $O/s30_dram_steady.csv
echo done
