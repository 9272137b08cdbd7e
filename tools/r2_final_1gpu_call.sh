#!/bin/bash
# Final 1-GPU check of a round: what the driver runs (suite with -x, smoke, both bench arms with default-ish flags).
set +e
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${T:-400} "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -2 gpurun_out/$name.log | cut -c1-600; }
T=700 run f_tests python -m pytest tests -x -q -m gpu -p no:cacheprovider
run f_smoke python __graft_entry__.py --smoke
run f_ref   python bench.py --impl reference --gpus 1 --steps 20 --warmup 5
run f_bench python bench.py --gpus 1 --steps 20 --warmup 5
run f_heads python tools/bench_heads.py
echo done
