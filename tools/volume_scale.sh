#!/bin/bash
N=${1:-2}; V=${2:-8}
timeout 600 python tools/volume_demo.py --volumes $V 2>&1 | grep VOLUME
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 tools/volume_demo.py --volumes $V 2>&1 | grep VOLUME
