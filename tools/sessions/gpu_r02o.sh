#!/bin/bash
# Two-GPU session: real multi-rank slab path (CUDA IPC peer regions) after the carried-sums rework, FT sharding, 2-GPU bench;
# and (one GPU) FT lanes 1/2/4 at the 8-GPU per-rank batch.
tag=${1:-r02_o}
o=gpurun_out
python -m pytest tests/test_multi_gpu.py -q 2>&1 | tail -30 > $o/${tag}_tests_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29555"
$TR --nproc-per-node 2 bench.py --gpus 2 --slab-only > $o/${tag}_slab2.json 2> $o/${tag}_slab2.err
$TR --nproc-per-node 2 bench.py --gpus 2 --steps 5 --warmup 3 > $o/${tag}_bench_2gpu.json 2> $o/${tag}_bench_2gpu.err
for l in 1 2 4; do echo "U1FT_LANES=$l" >> $o/${tag}_lanes128.log; U1FT_LANES=$l python tools/ft_small_batch.py 128 >> $o/${tag}_lanes128.log 2>&1; done
tail -n 5 $o/${tag}_tests_2gpu.log; grep -o '"slab": {"metric[^}]*"value": [0-9.e+]*' $o/${tag}_slab2.json $o/${tag}_bench_2gpu.json | sed 's/"metric.*"value"//'
cat $o/${tag}_lanes128.log; tail -c 300 $o/${tag}_bench_2gpu.err
