#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "stem" > gpurun_out/stem_test.txt 2>&1; echo "stem rc=$?"; tail -25 gpurun_out/stem_test.txt
