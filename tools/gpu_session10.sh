#!/bin/bash
# Round-2 GPU session 10: host-entry chunk schedule (count x ratio), C5 e2e.
O=gpurun_out
mkdir -p $O
B="python bench.py --no-others --no-cpu-baseline --no-parity --steps 20 --e2e-steps 100"
for c in 3 4 5; do for r in 1.5 2 3; do
  ANNF_HOST_CHUNKS=$c ANNF_HOST_CHUNK_RATIO=$r timeout -k 10 300 $B > $O/s10_e2e_c${c}_r${r}.json 2> $O/s10_e2e_c${c}_r${r}.err
done; done
echo done
