#!/bin/bash
# After tools/gpu_evidence.sh: copies the bench lines / logs from gpurun_out/ into profiles/ and regenerates the ncu summaries.
set -e
cd "$(dirname "$0")/.."
O=gpurun_out
for f in default reference c2 c3 c4 c5_tf32x3 k20_1 k20_2 k20_3; do cp $O/ev_bench_$f.json profiles/r02_bench_$f.json; done
cp $O/ev_pytest.log profiles/r02_pytest_gpu.log
cp $O/ev_phases.txt profiles/r02_c5_phases_final.txt
cp $O/ev_prep_flavours.txt profiles/r02_prep_flavours.txt
cp $O/ev_two_stream.txt profiles/r02_two_stream_probe.txt
python profiles/summarize.py launch $O/r02_launches_c5.csv profiles/r02_c5_launches.md
python profiles/summarize.py full $O/r02_prof_c5_step_final.ncu-rep profiles/r02_c5_step_kernels_summary.md \
    "prep_bulk_kernel=prep_kernel" "head_kernel=head_kernel" "(4, 8, 4)=gemm[1024x512x4500] fwd0 f16x3" "(36, 4, 1)=gemm[1024x4500x512] bwd0 f16x3"
python profiles/summarize.py full $O/r02_prof_c4_dmma.ncu-rep profiles/r02_c4_dmma_kernel_summary.md "annf_dmma_kernel=annf_dmma_kernel<f64>"
python profiles/sass_counts.py > profiles/r02_sass_counts.txt
