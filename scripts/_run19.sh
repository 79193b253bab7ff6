timeout 900 python -m pytest tests/test_text_encoders_gpu.py -m gpu -x -q -s 2>&1 | grep -E "kernel|passed|failed|Error|error|assert" | head -40
for v in 1 10 0; do echo variant $v; B2F_ATTN_VARIANT=$v timeout 100 python scripts/bench_kernels.py --what attn 2>&1 | grep -o "\"S\": [0-9]*, \"b2f_ms\": [0-9.]*, \"b2f_tflops\": [0-9.]*"; done
