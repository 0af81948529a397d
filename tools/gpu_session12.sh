#!/bin/bash
# Round-2 GPU session 12: DMMA weight-ring geometry (stages x KB) on C3 / C4: phase clocks of CTA 0 and us per call.
O=gpurun_out
mkdir -p $O
for v in 2x64 4x32 6x32 8x16; do
  echo "== ring $v"; ANNF_B200_LIB=$PWD/ann_force_b200/libannforce_b200_dmma_$v.so timeout -k 10 200 python tools/stream_phases.py time 2>&1 | tee $O/s12_dmma_$v.txt | tail -4
done
echo done
