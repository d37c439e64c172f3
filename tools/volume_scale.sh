#!/bin/bash
# volume prediction (BASELINE configs[2] shape: 155 slices x 256^2 per volume) at 1 and N GPUs, graph and eager
N=${1:-2}
V=${2:-4}
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -k "volume_graph" 2>&1 | tail -2
timeout 600 python tools/volume_demo.py --volumes $V 2>&1 | grep VOLUME
timeout 600 python tools/volume_demo.py --volumes $V --eager 2>&1 | grep VOLUME
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/volume_demo.py --volumes $V 2>&1 | grep VOLUME
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 tools/volume_demo.py --volumes $V --eager 2>&1 | grep VOLUME
