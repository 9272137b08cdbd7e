#!/bin/bash
# 1-GPU visit: item order A/B of the S-kernels (ONEPROT_SC: row chunks per chunk group; 16 = the old column-major order
# at N = 32768), the HBM-bound row kernels after their rewrite, the whole GPU suite.
set +e
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout ${T:-300} "$@" > gpurun_out/$name.log 2>&1; echo "exit $? ($name)"; tail -2 gpurun_out/$name.log | cut -c1-250; }
for sc in 16 8 4 2 1; do
  ONEPROT_SC=$sc run sc${sc}_fwd_e python tools/run_kernel.py fwd_e 32768 32768 1024 10
  ONEPROT_SC=$sc run sc${sc}_fwd   python tools/run_kernel.py fwd 32768 32768 1024 10
  ONEPROT_SC=$sc run sc${sc}_dz    python tools/run_kernel.py dz 16384 32768 1024 10
done
for sc in 16 4 2; do
  ONEPROT_SC=$sc run sc${sc}_bench python bench.py --steps 100 --warmup 5 --no-cpu-baseline
done
run heads_bench python tools/bench_heads.py
T=600 run t_all python -m pytest tests -q -m gpu -p no:cacheprovider -x
echo done
