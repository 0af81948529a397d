#!/bin/bash
# weak-scaling check with the driver's own command line: N ranks, --steps 20 --warmup 5
N=$1
O=gpurun_out
mkdir -p $O
for rep in 1 2; do
  timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + rep)) bench.py --gpus $N --steps 20 --warmup 5 > $O/scale_n${N}_$rep.json 2> $O/scale_n${N}_$rep.err
  tail -c 400 $O/scale_n${N}_$rep.json | head -c 400; echo
done
if [ "$N" = "8" ]; then
  for n in 1 2 4; do
    if [ $n = 1 ]; then timeout -k 10 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-others > $O/scale_n1_of8box.json 2> $O/scale_n1_of8box.err
    else timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps 20 --warmup 5 > $O/scale_n${n}_of8box.json 2> $O/scale_n${n}_of8box.err; fi
  done
fi
echo done
