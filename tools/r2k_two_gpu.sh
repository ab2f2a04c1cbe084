timeout 400 python -m pytest tests/test_gpu_shard.py -m gpu -q 2>&1 | tail -15
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 8 --warmup 3 --no-sub-records > gpurun_out/r2k_bench_n2.json 2> gpurun_out/r2k_bench_n2.err; echo "n2 rc=$?"
tail -c 500 gpurun_out/r2k_bench_n2.err; cut -c1-200 gpurun_out/r2k_bench_n2.json
