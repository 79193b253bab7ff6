scripts/gpu_tests.sh tests/test_cli_gpu.py tests/test_text_encoders_gpu.py tests/test_pipeline_gpu.py 2>&1 | grep -E "^==|passed|failed|Error" | head -20
echo BENCH; timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo rc=$?; cat gpurun_out/bench_full.json | head -c 1500; tail -5 gpurun_out/bench_full.err
