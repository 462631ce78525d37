set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r1_bench_final2.json 2> gpurun_out/bench_final2.err; tail -c 600 gpurun_out/r1_bench_final2.json
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/r1_f7_launches.csv python profiles/run_update.py 2 > gpurun_out/ncu_f7.log 2>&1; tail -2 gpurun_out/ncu_f7.log
ncu --set full --clock-control none --import-source on -k regex:mlp_fwd_tc_kernel -s 6 -c 1 -o gpurun_out/r1_f7_mlp_fwd python profiles/run_update.py 2 > gpurun_out/ncu_f7b.log 2>&1; tail -2 gpurun_out/ncu_f7b.log
ls -la gpurun_out | tail -5
