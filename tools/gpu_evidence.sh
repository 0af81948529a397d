#!/bin/bash
# Round-2 evidence run: everything bench.py / the tests / ncu produce for the committed profiles/ summaries.
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.draw,temperature.gpu --format=csv > $O/ev_smi.txt 2>&1
echo "== pytest gpu"; timeout -k 10 1200 python -m pytest tests -q -m gpu --timeout 300 > $O/ev_pytest.log 2>&1; echo "rc=$?" >> $O/ev_pytest.log; tail -3 $O/ev_pytest.log
echo "== smoke"; timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/ev_smoke.log 2>&1; echo "rc=$?" >> $O/ev_smoke.log; tail -2 $O/ev_smoke.log
echo "== reference arm"; timeout -k 10 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/ev_bench_reference.json 2> $O/ev_bench_reference.err
for i in 1 2 3; do echo "== bench k20 #$i"; timeout -k 10 600 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/ev_bench_k20_$i.json 2> $O/ev_bench_k20_$i.err; done
echo "== bench default"; timeout -k 10 600 python bench.py > $O/ev_bench_default.json 2> $O/ev_bench_default.err
for w in c2 c3 c4; do echo "== bench $w"; timeout -k 10 300 python bench.py --workload $w > $O/ev_bench_$w.json 2> $O/ev_bench_$w.err; done
echo "== bench c5 tf32x3"; timeout -k 10 300 python bench.py --precision tf32x3 --no-others --no-cpu-baseline > $O/ev_bench_c5_tf32x3.json 2> $O/ev_bench_c5_tf32x3.err
echo "== prep flavours"
for spec in "0 1" "0 2" "0 4" "0 8" "1 4" "2 4"; do set -- $spec
  echo "mode $1 warps $2: $(ANNF_PREP_MODE=$1 ANNF_PREP_WARPS=$2 timeout -k 10 300 python bench.py --steps 20 --warmup 5 --no-others --no-cpu-baseline --no-parity 2>/dev/null | python -c 'import json,sys; print(json.load(sys.stdin)["ms_per_step"])')"
done > $O/ev_prep_flavours.txt 2>&1; cat $O/ev_prep_flavours.txt
echo "== two streams"; timeout -k 10 300 python tools/two_stream_probe.py --replicas 1024 > $O/ev_two_stream.txt 2>&1; timeout -k 10 300 python tools/two_stream_probe.py --replicas 2048 >> $O/ev_two_stream.txt 2>&1; cat $O/ev_two_stream.txt
echo "== launch list"
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_c5.csv python bench.py --no-others --no-cpu-baseline --no-parity --steps 20 --warmup 3 --graph off --e2e-steps 1 > $O/ev_ncu_launch.log 2>&1
echo "== ncu full (one step)"
timeout -k 10 900 ncu --set full --clock-control none --import-source on --launch-skip 120 --launch-count 6 -o $O/r02_prof_c5_step_final -f \
   python bench.py --no-others --no-cpu-baseline --no-parity --steps 20 --warmup 3 --graph off --e2e-steps 1 > $O/ev_ncu_full.log 2>&1; tail -2 $O/ev_ncu_full.log
echo "== ncu full c4 / c2"
timeout -k 10 600 ncu --set full --clock-control none --import-source on --launch-skip 50 --launch-count 1 -o $O/r02_prof_c4_dmma -f \
   python bench.py --workload c4 --no-others --no-cpu-baseline --no-parity --steps 20 --warmup 3 --graph off --e2e-steps 1 > $O/ev_ncu_c4.log 2>&1
export ANNF_B200_LIB=$PWD/ann_force_b200/libannforce_b200_clocks.so
echo "== phases"; timeout -k 10 300 python tools/tensor_phases.py c5 > $O/ev_phases.txt 2>&1; tail -13 $O/ev_phases.txt | head -12
echo done
