#!/bin/bash
# Short multi-GPU check: the multi-rank parity tests only (+ one bench line at that GPU count).
set +e
mkdir -p gpurun_out
G=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_zb_configs.py tests/test_gpu_multirank.py tests/test_gpu_z2_sequencer.py tests/test_gpu_z4_siglip.py tests/test_gpu_za_keep_exp.py -q -m gpu -p no:cacheprovider -k "zb_configs or multi_gpu or two_gpu" > gpurun_out/c_tests.log 2>&1; echo "tests exit $?"; tail -3 gpurun_out/c_tests.log
timeout 300 $TR --master-port 29541 bench.py --gpus $G --steps 100 --warmup 5 > gpurun_out/c_bench.log 2>&1; echo "bench exit $?"
echo done
