#!/bin/bash
O=gpurun_out
mkdir -p $O
for cfg in "3 1.5" "4 1" "4 1.5" "2 1"; do set -- $cfg
  echo "== chunks $1 ratio $2"; ANNF_HOST_TRACE=1 ANNF_HOST_CHUNKS=$1 ANNF_HOST_CHUNK_RATIO=$2 timeout -k 10 200 python tools/host_trace.py 2>&1 | tail -2 | tee $O/s11_trace_c$1_r$2.txt
done
echo done
