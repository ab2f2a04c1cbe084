# 2-GPU validation of the sharded path: parity tests, then the bench at N=2 (driver-style torchrun launch) and N=1 for the same build
nvidia-smi -L
python -m pytest tests/test_gpu_shard.py tests/test_gpu_pusch.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2e_tests.txt; tail -5 gpurun_out/r2e_tests.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 --no-sub-records > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err; echo "n2 rc=$?"
timeout 600 python bench.py --gpus 1 --steps 8 --warmup 3 --no-sub-records --no-cpu-baseline > gpurun_out/r2e_bench_n1.json 2> gpurun_out/r2e_bench_n1.err; echo "n1 rc=$?"
tail -c 1500 gpurun_out/r2e_bench_n2.err; cat gpurun_out/r2e_bench_n2.json | cut -c1-1500
