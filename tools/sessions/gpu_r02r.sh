#!/bin/bash
# 8-GPU session on the final code: multi-rank slab and FT-sharding checks at 8 ranks, full bench line, slab-only line.
tag=${1:-r02_r}
o=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29555"
( $TR --nproc-per-node 8 tests/slab_worker.py 256 4 peer 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -3
  $TR --nproc-per-node 8 tests/slab_worker.py 512 16 peer 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -3
  $TR --nproc-per-node 8 tests/slab_worker.py 256 4 nccl 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -3
  $TR --nproc-per-node 8 tests/ft_shard_worker.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -3 ) > $o/${tag}_workers_8gpu.log 2>&1
$TR --nproc-per-node 8 bench.py --gpus 8 --steps 5 --warmup 3 > $o/${tag}_bench_8gpu.json 2> $o/${tag}_bench_8gpu.err
$TR --nproc-per-node 8 bench.py --gpus 8 --slab-only > $o/${tag}_slab8.json 2> /dev/null
$TR --nproc-per-node 4 bench.py --gpus 4 --slab-only > $o/${tag}_slab4.json 2> /dev/null
cat $o/${tag}_workers_8gpu.log
grep -h -o '"slab": {"metric[^}]*"value": [0-9.e+]*' $o/${tag}_slab8.json $o/${tag}_slab4.json $o/${tag}_bench_8gpu.json | sed 's/"metric.*"value"//'
tail -c 300 $o/${tag}_bench_8gpu.err
