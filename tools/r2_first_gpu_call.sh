#!/bin/bash
# First gpurun call of the next round: everything that was built after round 1's GPU budget was
# spent, in one box visit (1 GPU).  Each step writes its own log under gpurun_out/ and never stops the
# script, so that one failure does not hide the rest.
#     gpurun --timeout 2400 -- 'bash tools/r2_first_gpu_call.sh'
set +e
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${T:-420} "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -3 gpurun_out/$name.log; }
run t_validated   python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_multirank.py -k "not _z"
for f in z1_prefetch z2_sequencer z3_heads z4_siglip z5_dz_l2_hints z6_retrieval z7_robust z8_graph z9_module_replica za_keep_exp; do
  run t_$f python -m pytest tests/test_gpu_$f.py -q -m gpu
done
run smoke         python __graft_entry__.py --smoke
run bench_default python bench.py --steps 10 --warmup 3
ONEPROT_SEQ=1 run bench_seq python bench.py --steps 10 --warmup 3 --no-cpu-baseline
ONEPROT_KEEP_EXP=1 run bench_keep_exp python bench.py --steps 10 --warmup 3 --no-cpu-baseline
ONEPROT_KEEP_EXP=1 ONEPROT_KEEP_OVERLAP=1 run bench_keep_overlap2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
ONEPROT_KEEP_EXP=1 ONEPROT_KEEP_OVERLAP=1 ONEPROT_KEEP_PANELS=4 run bench_keep_overlap4 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
ONEPROT_KEEP_EXP=1 ONEPROT_KEEP_OVERLAP=1 ONEPROT_KEEP_PANELS=8 run bench_keep_overlap8 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run dz_default    python tools/run_kernel.py dz 16384 32768 1024 10
run k_fwd         python tools/run_kernel.py fwd 32768 32768 1024 10
run k_fwd_e       python tools/run_kernel.py fwd_e 32768 32768 1024 10
run k_dz_e        python tools/run_kernel.py dz_e 32768 32768 1024 10
ONEPROT_DZ_L2_HINTS=1 run k_fwd_e_l2 python tools/run_kernel.py fwd_e 32768 32768 1024 10
ONEPROT_DZ_L2_HINTS=1 run dz_l2_hints python tools/run_kernel.py dz 16384 32768 1024 10
ONEPROT_DZ_L2_HINTS=1 run bench_dz_l2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run host_1024     python tools/host_overhead.py 1024
run host_4096     python tools/host_overhead.py 4096
run eager_bar     python tests/perf_eager_bar.py --sizes 8192,32768 --reps 5
run heads_bench   python tools/bench_heads.py
grep -h '"metric"' gpurun_out/bench_default.log gpurun_out/bench_seq.log gpurun_out/bench_keep_exp.log gpurun_out/bench_keep_overlap2.log gpurun_out/bench_keep_overlap4.log gpurun_out/bench_keep_overlap8.log > gpurun_out/r2_bench_lines.json
echo done
echo "python tools/show_bench.py -v gpurun_out/r2_bench_lines.json   # side-by-side view of the A/B lines"
