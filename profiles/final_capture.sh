# Round-end capture: full bench line, GPU tests, the larger-batch point, then the ncu launch list of two updates.
set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r1_bench_final3.json 2> gpurun_out/bench_final3.err; tail -c 300 gpurun_out/r1_bench_final3.json
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 --batch 2048 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('B2048', round(d['value']), d['ms_per_step'], round(d['e2e']['value']))"
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --csv --log-file gpurun_out/r1_f8_launches.csv python profiles/run_update.py 2 > gpurun_out/ncu_f8.log 2>&1; tail -1 gpurun_out/ncu_f8.log
