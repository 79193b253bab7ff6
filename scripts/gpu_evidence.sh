#!/bin/bash
# One GPU-box pass that produces the evidence files of a round (copy what you want judged from gpurun_out/ to profiles/):
#   ncu --set full captures of the inference kernels (one double + one single block at C1024) and of the training kernels
#   (one double + one single block at the 512x512 training shapes), the ncu launch list of the bench command itself, the CPU
#   reference arm with its once-only extras, and the default bench lines (edit, train512).
# usage: bash scripts/gpu_evidence.sh [tag]      (tag names the output files, default r02)
set -u
cd "$(dirname "$0")/.."
TAG=${1:-r02}
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# (the .ncu-rep files are summarised on the box and removed: gpurun brings back at most 64 MiB)
$NCU --set full -k regex:'attn_fwd_kernel|gemm_bf16_2cta|ln_modulate_kernel' -c 30 -f -o /tmp/${TAG}_full_infer \
    python scripts/profile_step.py --layers 1,1 --forwards 1 > gpurun_out/${TAG}_full_infer.log 2>&1
echo "ncu infer rc=$?"
python scripts/ncu_summary.py /tmp/${TAG}_full_infer.ncu-rep > gpurun_out/${TAG}_ncu_full_infer_summary.json
$NCU --set full -k regex:'attn_bwd_kernel|gemm_grad_2cta|ln_modulate_bwd|rmsnorm_rope_bwd|adamw|attn_fwd_kernel' -c 30 -f \
    -o /tmp/${TAG}_full_train python scripts/profile_train.py --layers 1,1 > gpurun_out/${TAG}_full_train.log 2>&1
echo "ncu train rc=$?"
python scripts/ncu_summary.py /tmp/${TAG}_full_train.ncu-rep > gpurun_out/${TAG}_ncu_full_train_summary.json
$NCU --metrics gpu__time_duration.sum -c 9000 --csv --log-file /tmp/${TAG}_launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1
echo "ncu launch list rc=$?"
python scripts/ncu_launch_summary.py /tmp/${TAG}_launches_bench.csv "ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 python bench.py --steps 1 --warmup 1 (first 9000 launches: conditioning, VAE encode, 28 denoising steps)" > gpurun_out/${TAG}_ncu_launch_summary_bench.json
python bench.py --impl reference --cpu-extras --steps 1 --warmup 0 > gpurun_out/${TAG}_cpu_reference_extras.json 2> gpurun_out/${TAG}_cpu_reference_extras.err
echo "cpu reference rc=$?"
python bench.py > gpurun_out/${TAG}_bench_c1024.json 2> gpurun_out/${TAG}_bench_c1024.err
echo "bench rc=$?"; tail -c 400 gpurun_out/${TAG}_bench_c1024.json
python bench.py --workload train512 > gpurun_out/${TAG}_bench_train512.json 2> gpurun_out/${TAG}_bench_train512.err
echo "bench train rc=$?"; tail -c 300 gpurun_out/${TAG}_bench_train512.json
du -sh gpurun_out | tail -1
