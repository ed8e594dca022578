#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_train.py -m gpu -q -x -k "batchnorm_train" 2>&1 | tail -8
