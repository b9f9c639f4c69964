#!/bin/bash
# Multi-rank check of the FINAL code on N = $2 GPUs: slab workers (peer, nccl), FT sharding worker, slab-only bench line.
tag=${1:-r02_f2}; N=${2:-2}
o=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29555"
( $TR --nproc-per-node $N tests/slab_worker.py 512 16 peer 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -2
  $TR --nproc-per-node $N tests/slab_worker.py 256 4 nccl 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -2
  $TR --nproc-per-node $N tests/ft_shard_worker.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -2 ) > $o/${tag}_workers.log 2>&1
$TR --nproc-per-node $N bench.py --gpus $N --slab-only > $o/${tag}_slab.json 2> /dev/null
cat $o/${tag}_workers.log
grep -h -o '"slab": {"metric[^}]*"value": [0-9.e+]*\|"ms_per_step": [0-9.]*\|"ok": [a-z]*' $o/${tag}_slab.json | sed 's/"metric.*"value"//'
