#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "fused_groupnorm_operand" > gpurun_out/xf_test.txt 2>&1; echo "xf rc=$?"; tail -30 gpurun_out/xf_test.txt
