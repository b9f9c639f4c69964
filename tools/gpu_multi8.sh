#!/bin/bash
# 8-GPU session: multi-rank tests (2/4/8 ranks), full bench at 8 ranks, then slab-only sweeps of the halo width.
tag=${1:-r02_d}
python -m pytest tests/test_multi_gpu.py -q 2>&1 | tail -40 > gpurun_out/${tag}_tests_8gpu.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29555"
$TR --nproc-per-node 8 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/${tag}_bench_8gpu.json 2> gpurun_out/${tag}_bench_8gpu.err
for k in 8 12 20 24 32; do
  U1_SLAB_HALO=$k $TR --nproc-per-node 8 bench.py --gpus 8 --slab-only > gpurun_out/${tag}_slab8_k$k.json 2> gpurun_out/${tag}_slab8_k$k.err
done
U1_MULTISTEP_EW=64 U1_SLAB_HALO=24 $TR --nproc-per-node 8 bench.py --gpus 8 --slab-only > gpurun_out/${tag}_slab8_k24_ew64.json 2> /dev/null
U1_SLAB_TRANSPORT=nccl $TR --nproc-per-node 8 bench.py --gpus 8 --slab-only > gpurun_out/${tag}_slab8_nccl.json 2> /dev/null
$TR --nproc-per-node 4 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/${tag}_bench_4gpu.json 2> gpurun_out/${tag}_bench_4gpu.err
tail -6 gpurun_out/${tag}_tests_8gpu.log
grep -h -o '"slab": {"metric[^}]*"value": [0-9.e+]*' gpurun_out/${tag}_slab8_*.json gpurun_out/${tag}_bench_8gpu.json | sed 's/"metric.*"value"//'
