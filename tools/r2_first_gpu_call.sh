#!/bin/bash
# First gpurun call of round 2 (1 GPU): the whole GPU suite WITHOUT -x, the prepared A/B bench lines in the sustained
# regime (100 timed steps), the isolated kernels, the eager-PyTorch-on-B200 bar and the head-kernel HBM roofline.
# Each step writes its own log under gpurun_out/ and never stops the script.
#     gpurun --timeout 1500 -- 'bash tools/r2_first_gpu_call.sh'
set +e
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${T:-300} "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -3 gpurun_out/$name.log; }
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm --format=csv > gpurun_out/smi.log 2>&1
T=600 run t_all python -m pytest tests -q -m gpu -s -p no:cacheprovider
run smoke         python __graft_entry__.py --smoke
S="--steps 100 --warmup 5"
run bench_default python bench.py $S
ONEPROT_SEQ=1 run bench_seq python bench.py $S --no-cpu-baseline
ONEPROT_KEEP_EXP=1 run bench_keep_exp python bench.py $S --no-cpu-baseline
ONEPROT_KEEP_EXP=1 ONEPROT_KEEP_OVERLAP=1 run bench_keep_overlap2 python bench.py $S --no-cpu-baseline
ONEPROT_KEEP_EXP=1 ONEPROT_KEEP_OVERLAP=1 ONEPROT_KEEP_PANELS=4 run bench_keep_overlap4 python bench.py $S --no-cpu-baseline
ONEPROT_KEEP_EXP=1 ONEPROT_KEEP_OVERLAP=1 ONEPROT_KEEP_PANELS=8 run bench_keep_overlap8 python bench.py $S --no-cpu-baseline
ONEPROT_KEEP_EXP=1 ONEPROT_SEQ=1 run bench_keep_seq python bench.py $S --no-cpu-baseline
ONEPROT_DZ_L2_HINTS=1 run bench_dz_l2 python bench.py $S --no-cpu-baseline
ONEPROT_CG2=1 run bench_cg2 python bench.py $S --no-cpu-baseline
run dz_default    python tools/run_kernel.py dz 16384 32768 1024 10
run k_fwd         python tools/run_kernel.py fwd 32768 32768 1024 10
run k_fwd_e       python tools/run_kernel.py fwd_e 32768 32768 1024 10
run k_dz_e        python tools/run_kernel.py dz_e 32768 32768 1024 10
ONEPROT_DZ_L2_HINTS=1 run k_fwd_e_l2 python tools/run_kernel.py fwd_e 32768 32768 1024 10
ONEPROT_DZ_L2_HINTS=1 run dz_l2_hints python tools/run_kernel.py dz 16384 32768 1024 10
run host_1024     python tools/host_overhead.py 1024
run host_4096     python tools/host_overhead.py 4096
run eager_bar     python tests/perf_eager_bar.py --sizes 8192,32768 --reps 5
run heads_bench   python tools/bench_heads.py
grep -h '"metric"' gpurun_out/bench_default.log gpurun_out/bench_seq.log gpurun_out/bench_keep_exp.log gpurun_out/bench_keep_overlap2.log gpurun_out/bench_keep_overlap4.log gpurun_out/bench_keep_overlap8.log gpurun_out/bench_keep_seq.log gpurun_out/bench_dz_l2.log gpurun_out/bench_cg2.log > gpurun_out/r2_bench_lines.json
echo done
