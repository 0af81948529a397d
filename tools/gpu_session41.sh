#!/bin/bash
# Round-2 GPU session 41: L2 hint masks again, after the prologue fix (default library = mask 7).
O=gpurun_out
mkdir -p $O
run() { timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-others --no-cpu-baseline --no-parity 2>/dev/null | python -c 'import json,sys; print(json.load(sys.stdin)["ms_per_step"])'; }
for rep in 1 2; do
echo "mask 7 (default lib): $(run)"
for h in 0 3 15; do
  echo "mask $h: $(ANNF_B200_LIB=$PWD/ann_force_b200/_variants/libannforce_b200_h$h.so run)"
done
done > $O/s41_hints.txt 2>&1; cat $O/s41_hints.txt
echo done
