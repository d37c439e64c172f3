#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "volume" > gpurun_out/vol_test.txt 2>&1; echo "vol rc=$?"; tail -40 gpurun_out/vol_test.txt
