python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2h_tests.txt; tail -3 gpurun_out/r2h_tests.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/r2h_ncu_bench.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 --no-sub-records > gpurun_out/r2h_bench_n2.json 2> gpurun_out/r2h_bench_n2.err; echo "n2 rc=$?"
timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 --no-sub-records --no-cpu-baseline > gpurun_out/r2h_bench_n1.json 2> gpurun_out/r2h_bench_n1.err; echo "n1 rc=$?"
cut -c1-300 gpurun_out/r2h_bench_n1.json; cut -c1-300 gpurun_out/r2h_bench_n2.json
