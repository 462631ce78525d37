# Round-2 ncu capture (run under gpurun, one GPU): the launch list of bench.py itself, then `--set full` of the dominant
# kernel (fused MLP forward) and of the long-sequence attention forward.  Every ncu run follows a plain run of the
# same command that exited 0.
set -x
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sustained --no-extras"
$BENCH > gpurun_out/r2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches.csv $BENCH > gpurun_out/r2_ncu1.log 2>&1
$BENCH > gpurun_out/r2_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:mlp_fwd_tc_kernel -s 40 -c 3 -o gpurun_out/r2_prof_mlp $BENCH > gpurun_out/r2_ncu2.log 2>&1
python profiles/c5_breakdown.py 64 > gpurun_out/r2_plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_.*long -s 20 -c 6 -o gpurun_out/r2_prof_attnl python profiles/c5_breakdown.py 64 > gpurun_out/r2_ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
