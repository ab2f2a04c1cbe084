python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2p_tests.txt; tail -4 gpurun_out/r2p_tests.txt
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo bench rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2p_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/r2p_ncu_bench.log 2>&1
cut -c1-260 gpurun_out/r2p_bench.json
