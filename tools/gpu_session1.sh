#!/bin/bash
# Round-2 GPU session 1: parity of the f16x3 tensor path, harness reproducibility, chunk-length / tile-shape experiments.
# Run through gpurun from the repo root; everything lands in gpurun_out/s1_*.
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/s1_smi.txt 2>&1
echo "== pytest gpu" ; timeout 900 python -m pytest tests -q -m gpu --timeout 600 -x > $O/s1_pytest.log 2>&1; echo "rc=$?" >> $O/s1_pytest.log; tail -5 $O/s1_pytest.log
echo "== smoke" ; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/s1_smoke.log 2>&1; echo "rc=$?" >> $O/s1_smoke.log; tail -2 $O/s1_smoke.log
B="python bench.py --no-others --no-cpu-baseline"
for i in 1 2 3; do echo "== bench steps20 #$i"; timeout 300 $B --steps 20 --warmup 5 > $O/s1_bench_k20_$i.json 2> $O/s1_bench_k20_$i.err; done
echo "== bench default"; timeout 600 python bench.py > $O/s1_bench_default.json 2> $O/s1_bench_default.err
echo "== bench tf32x3"; ANNF_AUTO_TENSOR=tf32x3 timeout 300 $B --steps 200 > $O/s1_bench_tf32.json 2> $O/s1_bench_tf32.err
for c in 1 4; do
  echo "== chunk $c f16"; ANNF_TENSOR_CHUNK=$c timeout 300 $B --steps 200 > $O/s1_bench_f16_chunk$c.json 2> $O/s1_bench_f16_chunk$c.err
  echo "== chunk $c tf32"; ANNF_AUTO_TENSOR=tf32x3 ANNF_TENSOR_CHUNK=$c timeout 300 $B --steps 200 > $O/s1_bench_tf32_chunk$c.json 2> $O/s1_bench_tf32_chunk$c.err
done
for bn in 128 130 132 256 512; do
  echo "== f16 bn $bn"; ANNF_GEMM_BN=$bn timeout 300 $B --steps 200 > $O/s1_bench_f16_bn$bn.json 2> $O/s1_bench_f16_bn$bn.err
done
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/s1_bench_reference.json 2> $O/s1_bench_reference.err
echo "== launch list (ncu, one window)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/s1_launches_c5.csv python bench.py --no-others --no-cpu-baseline --no-parity --steps 20 --warmup 3 --graph off > $O/s1_ncu_launch.log 2>&1
echo done
