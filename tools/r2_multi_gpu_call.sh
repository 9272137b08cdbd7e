#!/bin/bash
# Multi-GPU visit (gpurun --gpus 2 or 8): every multi-rank parity test, then bench.py at that GPU count with the Python
# host and the C step sequencer, the multi-rank eager bar, and (8 GPUs) the scaling pair 1 vs N on the same box.
#     gpurun --gpus 8 --timeout 900 -- 'bash tools/r2_multi_gpu_call.sh'
set +e
mkdir -p gpurun_out
G=$(nvidia-smi -L | wc -l)
run() { name=$1; shift; echo "=== $name"; timeout ${T:-300} "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -3 gpurun_out/$name.log; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
T=600 run mg_tests python -m pytest tests/test_gpu_zb_configs.py tests/test_gpu_multirank.py tests/test_gpu_z2_sequencer.py tests/test_gpu_z4_siglip.py tests/test_gpu_za_keep_exp.py -q -m gpu -s -p no:cacheprovider -k "zb_configs or multi_gpu or two_gpu"
S="--steps 100 --warmup 5"
ONEPROT_BENCH_HOST=python run mg_bench_py   $TR --master-port 29511 bench.py --gpus $G $S
run mg_bench_seq $TR --master-port 29512 bench.py --gpus $G $S
ONEPROT_BENCH_HOST=python run mg_bench_py2  $TR --master-port 29513 bench.py --gpus $G $S
run mg_bench_seq2 $TR --master-port 29514 bench.py --gpus $G $S
run mg_bench_1    python bench.py --gpus 1 $S --no-cpu-baseline
run mg_eager      $TR --master-port 29515 tests/perf_eager_bar.py --world --sizes 8192,32768 --reps 5 --out gpurun_out/eager_bar_w$G.json
grep -h '"metric"' gpurun_out/mg_bench_py.log gpurun_out/mg_bench_seq.log gpurun_out/mg_bench_py2.log gpurun_out/mg_bench_seq2.log gpurun_out/mg_bench_1.log > gpurun_out/r2_mg${G}_bench_lines.json
echo done
