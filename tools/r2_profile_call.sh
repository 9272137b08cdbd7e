#!/bin/bash
# ncu evidence for round 2 (one gpurun call, 1 GPU; B200_PROFILING.md: the plain run of the same command
# line comes first, every ncu run of a call counts as one).  Afterwards, here:
#     python profiles/summarize_ncu.py r2       # -> profiles/r2_launches_summary.txt, r2_ncu_raw_summary.csv, ncu_traffic.json
#     gpurun --timeout 1500 -- 'bash tools/r2_profile_call.sh'
set +e
mkdir -p gpurun_out
BENCH="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$BENCH > gpurun_out/r2_plain.log 2>&1 || { echo "plain bench run failed"; tail -5 gpurun_out/r2_plain.log; exit 1; }
# (1) every launch of a short bench run with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches.csv $BENCH \
    > gpurun_out/r2_ncu_launches.log 2>&1
echo "launch list exit $?"
# (2) full-set capture of the four tensor-core kernels of one step (skip the warm-up launches of each)
ncu --set full --clock-control none --import-source on -k regex:'clip_s_kernel|gemm_kernel' -s 24 -c 8 \
    -o gpurun_out/r2_prof -f $BENCH > gpurun_out/r2_ncu_full.log 2>&1
echo "full-set exit $?"
# (3) the dL/dZ kernel alone, default vs L2 hints (source-level view for the MMA-warp stall analysis)
python tools/run_kernel.py dz 16384 32768 1024 3 > gpurun_out/r2_dz_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:clip_s_kernel -s 1 -c 1 -o gpurun_out/r2_dz -f \
    python tools/run_kernel.py dz 16384 32768 1024 3 > gpurun_out/r2_ncu_dz.log 2>&1
ONEPROT_DZ_L2_HINTS=1 python tools/run_kernel.py dz 16384 32768 1024 3 > gpurun_out/r2_dz_l2_plain.log 2>&1 &&
ONEPROT_DZ_L2_HINTS=1 ncu --set full --clock-control none --import-source on -k regex:clip_s_kernel -s 1 -c 1 -o gpurun_out/r2_dz_l2 -f \
    python tools/run_kernel.py dz 16384 32768 1024 3 > gpurun_out/r2_ncu_dz_l2.log 2>&1
# (4) the stored-exponentials variant: launch list + full set of its kernels (FWD_E, dz_from_exp, the two whole-panel GEMMs)
ONEPROT_KEEP_EXP=1 $BENCH > gpurun_out/r2_keep_plain.log 2>&1 && {
ONEPROT_KEEP_EXP=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_keep_launches.csv $BENCH \
    > gpurun_out/r2_ncu_keep_launches.log 2>&1
ONEPROT_KEEP_EXP=1 ncu --set full --clock-control none --import-source on -k regex:'clip_s_kernel|gemm_kernel|dz_from_exp' -s 24 -c 8 \
    -o gpurun_out/r2_prof_keep -f $BENCH > gpurun_out/r2_ncu_keep_full.log 2>&1
echo "keep_exp captures exit $?"; }
ls -la gpurun_out/r2_*
echo done
