#!/bin/bash
# run on the GPU box: full-size tests + bench lines for every workload
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -5
