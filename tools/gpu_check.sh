#!/bin/bash
# One-GPU check run used with gpurun: GPU tests, smoke, a short bench and the ncu counter pass.
tag=${1:-r02_a}
python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/${tag}_tests.log
python tests/slab_worker.py 128 4 peer 2 > gpurun_out/${tag}_virt2.log 2>&1
python tests/slab_worker.py 512 16 peer 8 > gpurun_out/${tag}_virt8.log 2>&1
python tests/slab_worker.py 128 3 nccl > gpurun_out/${tag}_w1nccl.log 2>&1
python tests/ft_shard_worker.py > gpurun_out/${tag}_ftshard.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_fp64.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:"k_traj_resident|k_leapfrog_step" --csv --log-file gpurun_out/${tag}_counters.csv \
    python tools/prof_counters.py > gpurun_out/${tag}_counters.log 2>&1
tail -8 gpurun_out/${tag}_tests.log; tail -3 gpurun_out/${tag}_virt2.log gpurun_out/${tag}_virt8.log gpurun_out/${tag}_w1nccl.log gpurun_out/${tag}_ftshard.log
tail -c 800 gpurun_out/${tag}_bench.err
