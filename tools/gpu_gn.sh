#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "single_pass" > gpurun_out/gn_test.txt 2>&1; echo "gn rc=$?"; tail -25 gpurun_out/gn_test.txt
