#!/bin/bash
# Multi-GPU check used with `gpurun --gpus N`: multi-rank tests, then the bench at N ranks (and optionally fewer).
N=${1:-2}; tag=${2:-r02_m}; shift 2
python -m pytest tests/test_multi_gpu.py -q 2>&1 | tail -40 > gpurun_out/${tag}_tests_${N}gpu.log
for n in "$@" $N; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $n --steps 5 --warmup 3 \
      > gpurun_out/${tag}_bench_${n}gpu.json 2> gpurun_out/${tag}_bench_${n}gpu.err
done
tail -6 gpurun_out/${tag}_tests_${N}gpu.log
for n in "$@" $N; do tail -c 400 gpurun_out/${tag}_bench_${n}gpu.err; done
