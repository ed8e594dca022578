#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_f2_fixture.py tests/test_gpu_api.py -m gpu -q 2>&1 | tail -60 > gpurun_out/r2_pytest_b.txt
tail -8 gpurun_out/r2_pytest_b.txt
grep f2_bf16_train_grad gpurun_out/parity_notes.jsonl | tail -1 | cut -c1-1200
