set +e
mkdir -p gpurun_out
G=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1"
run() { name=$1; shift; echo "=== $name"; timeout ${T:-300} "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -2 gpurun_out/$name.log | cut -c1-1500; }
run m5_1gpu python bench.py --config modalities5 --steps 50 --warmup 5
run m5_ngpu $TR --master-port 29521 bench.py --gpus $G --config modalities5 --steps 50 --warmup 5
run e2e_small $TR --master-port 29522 tools/e2e_step.py --layers 4 --batch 128 --seq-len 64 --out gpurun_out/e2e_small.json
run e2e_1gpu python tools/e2e_step.py --batch 256 --seq-len 128 --out gpurun_out/e2e_1gpu.json
echo done
