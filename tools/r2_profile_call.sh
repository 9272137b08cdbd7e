#!/bin/bash
# ncu evidence for round 2 (one gpurun call, 1 GPU; B200_PROFILING.md: the plain run of the same command line comes
# first, a number printed under ncu is never a bench value).  Afterwards, here:
#     python profiles/summarize_ncu.py r2       # -> profiles/r2_launches_summary.txt, r2_ncu_raw_summary.csv, ncu_traffic.json
#     gpurun --timeout 1200 -- 'bash tools/r2_profile_call.sh'
set +e
mkdir -p gpurun_out
BENCH="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$BENCH > gpurun_out/r2_plain.log 2>&1 || { echo "plain bench run failed"; tail -5 gpurun_out/r2_plain.log; exit 1; }
# (1) every launch of a short bench run with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $BENCH \
    > gpurun_out/r2_ncu_launches.log 2>&1
echo "launch list exit $?"
# (2) full-set capture of the tensor-core kernels + the rescale of one step (skip the warm-up launches)
ncu --set full --clock-control none --import-source on -k regex:'clip_s_kernel|gemm_kernel|dz_from_exp' -s 16 -c 5 \
    -o gpurun_out/r2_prof -f $BENCH > gpurun_out/r2_ncu_full.log 2>&1
echo "full-set exit $?"
# (3) the kernels alone for the source-level stall analysis: FWD_E (dominant), FWD (same mainloop, no panel), dA GEMM
for k in fwd_e fwd da; do
  python tools/run_kernel.py $k 32768 32768 1024 3 > gpurun_out/r2_${k}_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:'clip_s_kernel|gemm_kernel' -c 4 \
      -o gpurun_out/r2_$k -f python tools/run_kernel.py $k 32768 32768 1024 3 > gpurun_out/r2_ncu_$k.log 2>&1
  echo "$k exit $?"
done
ls -la gpurun_out/r2_*
echo done
