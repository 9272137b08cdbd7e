#!/bin/bash
# Multi-GPU visit (gpurun --gpus 2 or 8): every multi-rank parity test, then bench.py at that GPU count with the C step
# sequencer (default) and the Python host, the scaling pair 1 vs N on the same box, BASELINE cfg 3 (modalities5) and
# cfg 5 (e2e step), the multi-rank eager bar, a kernel timeline of rank 0 and the NVLink byte counters around a bench run.
#     gpurun --gpus 8 --timeout 1200 -- 'bash tools/r2_multi_gpu_call.sh'
set +e
mkdir -p gpurun_out
G=$(nvidia-smi -L | wc -l)
run() { name=$1; shift; echo "=== $name"; timeout ${T:-300} "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -3 gpurun_out/$name.log | cut -c1-400; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
T=700 run mg_tests python -m pytest tests/test_gpu_zb_configs.py tests/test_gpu_multirank.py tests/test_gpu_z2_sequencer.py tests/test_gpu_z4_siglip.py tests/test_gpu_za_keep_exp.py -q -m gpu -s -p no:cacheprovider -k "zb_configs or multi_gpu or two_gpu"
S="--steps 100 --warmup 5"
nvidia-smi nvlink -gt d -i 0 > gpurun_out/nvlink_before.txt 2>&1
run mg_bench_seq  $TR --master-port 29512 bench.py --gpus $G $S
nvidia-smi nvlink -gt d -i 0 > gpurun_out/nvlink_after.txt 2>&1
ONEPROT_BENCH_HOST=python run mg_bench_py $TR --master-port 29511 bench.py --gpus $G $S
run mg_bench_seq2 $TR --master-port 29514 bench.py --gpus $G $S
run mg_bench_1    python bench.py --gpus 1 $S --no-cpu-baseline
run mg_m5         $TR --master-port 29516 bench.py --gpus $G --config modalities5 --steps 50 --warmup 5
run mg_eager      $TR --master-port 29515 tests/perf_eager_bar.py --world --sizes 8192,32768 --reps 5 --out gpurun_out/eager_bar_w$G.json
run mg_timeline   $TR --master-port 29517 tools/timeline.py
T=400 run mg_e2e  $TR --master-port 29518 tests/perf_e2e_step.py --batch 256 --seq-len 128 --out gpurun_out/e2e_w$G.json
grep -h '"metric"' gpurun_out/mg_bench_seq.log gpurun_out/mg_bench_py.log gpurun_out/mg_bench_seq2.log gpurun_out/mg_bench_1.log gpurun_out/mg_m5.log > gpurun_out/r2_mg${G}_bench_lines.json
echo done
