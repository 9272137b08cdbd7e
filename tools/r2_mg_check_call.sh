#!/bin/bash
# Short multi-GPU check: the multi-rank parity tests, two bench lines at that GPU count, BASELINE cfg 3 (modalities5).
set +e
mkdir -p gpurun_out
G=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_zb_configs.py tests/test_gpu_multirank.py tests/test_gpu_z2_sequencer.py tests/test_gpu_z4_siglip.py tests/test_gpu_z8_graph.py tests/test_gpu_za_keep_exp.py -q -m gpu -p no:cacheprovider -k "zb_configs or multi_gpu or two_gpu" > gpurun_out/c_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/c_tests.log
timeout 300 $TR --master-port 29541 bench.py --gpus $G --steps 20 --warmup 5 > gpurun_out/c_bench.log 2>&1; echo "bench exit $?"
timeout 300 $TR --master-port 29542 bench.py --gpus $G --steps 100 --warmup 5 > gpurun_out/c_bench100.log 2>&1; echo "bench100 exit $?"
timeout 300 $TR --master-port 29543 bench.py --gpus $G --config modalities5 --steps 50 --warmup 5 > gpurun_out/c_m5.log 2>&1; echo "m5 exit $?"
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/c_bench_1.log 2>&1; echo "bench1 exit $?"
grep -h '"metric"' gpurun_out/c_bench.log gpurun_out/c_bench100.log gpurun_out/c_m5.log gpurun_out/c_bench_1.log > gpurun_out/r2_final_mg${G}_lines.json
echo done
