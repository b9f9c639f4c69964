#!/bin/bash
# Final one-GPU checkpoint of a round: all GPU tests, smoke, full bench line, reference arm (short).
tag=${1:-r02_w}
o=gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -30 > $o/${tag}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $o/${tag}_smoke.log 2>&1
python bench.py --steps 10 --warmup 3 > $o/${tag}_bench.json 2> $o/${tag}_bench.err
tail -n 4 $o/${tag}_tests.log; tail -2 $o/${tag}_smoke.log; tail -c 400 $o/${tag}_bench.err
