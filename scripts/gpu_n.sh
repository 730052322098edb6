#!/bin/bash
# Multi-GPU iteration: the dist tests, then bench.py under torchrun on N GPUs.
#   gpurun --gpus N -- 'N=2 bash scripts/gpu_n.sh TAG [bench args]'
TAG=${1:-n}; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
N=${N:-2}
nvidia-smi topo -m > $OUT/topo.txt 2>&1
if [ "${SKIP_TESTS:-0}" != "1" ]; then
timeout 600 python -m pytest tests -q -m gpu -k "${K:-dist or onesided}" 2>&1 | tail -15 | tee $OUT/pytest.log
fi
timeout ${BENCH_TIMEOUT:-1500} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N "$@" 2> $OUT/bench.err | tail -2 | tee $OUT/bench_n$N.json | cut -c1-600
tail -15 $OUT/bench.err
