#!/bin/bash
# Final 1-GPU check of a round: what the driver runs (suite with -x, smoke, both bench arms), then the ncu evidence of the
# HBM-bound row kernels (dram bytes per launch next to the algorithmic bytes; plain run of the same command first).
set +e
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${T:-400} "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -2 gpurun_out/$name.log | cut -c1-600; }
T=700 run f_tests python -m pytest tests -x -q -m gpu -p no:cacheprovider
run f_smoke python __graft_entry__.py --smoke
run f_ref   python bench.py --impl reference --gpus 1 --steps 20 --warmup 5
run f_bench python bench.py --gpus 1 --steps 20 --warmup 5
run f_heads python tools/bench_heads.py
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none \
    -k regex:'layernorm|l2norm|gelu|meanpool|scale_kernel|dz_from_exp' -c 60 --csv --log-file gpurun_out/r2_rows_ncu.csv python tools/bench_heads.py > gpurun_out/f_heads_ncu.log 2>&1
echo "ncu rows exit $?"
echo done
