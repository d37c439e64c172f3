#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "attention" > gpurun_out/attn_test.txt 2>&1; echo "attn rc=$?"; tail -30 gpurun_out/attn_test.txt
