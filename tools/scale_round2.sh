# how the round-2 scaling record was produced (gpurun --gpus 8); see profiles/r2_scaling.md
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 8 --warmup 3 --no-sub-records > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "n8 rc=$?"
timeout 300 python -m pytest tests/test_gpu_shard.py -m gpu -q 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 4 --steps 8 --warmup 3 --no-sub-records > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err; echo "n4 rc=$?"
cut -c1-250 gpurun_out/r2_bench_n8.json; cut -c1-250 gpurun_out/r2_bench_n4.json
