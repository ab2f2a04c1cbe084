python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2t_tests.txt; tail -3 gpurun_out/r2t_tests.txt
python bench.py > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo bench rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2t_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/r2t_ncu_bench.log 2>&1
cut -c1-200 gpurun_out/r2t_bench.json
