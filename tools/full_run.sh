#!/bin/bash
# one gpurun call: the whole GPU suite, then the bench line.  usage: tools/full_run.sh <tag>
tag=${1:-run}
cd /root/repo
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu --durations=5 2>&1 | tail -15 ) > gpurun_out/${tag}_pytest.log 2>&1
( time python bench.py ) > gpurun_out/${tag}_bench.log 2>&1
tail -n 4 gpurun_out/${tag}_pytest.log; tail -n 6 gpurun_out/${tag}_bench.log | cut -c1-1500
