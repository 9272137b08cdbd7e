#!/bin/bash
# 1-GPU check of a kernel change: whole GPU suite, the isolated kernels, the bench line.
set +e
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${T:-400} "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -1 gpurun_out/$name.log | cut -c1-260; }
T=700 run k_tests python -m pytest tests -x -q -m gpu -p no:cacheprovider
for k in fwd_e fwd dz da db; do run k_$k python tools/run_kernel.py $k 32768 32768 1024 10; done
run k_bench python bench.py --steps 100 --warmup 5 --no-cpu-baseline
echo done
