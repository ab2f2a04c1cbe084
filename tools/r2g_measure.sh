python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2g_tests.txt; tail -4 gpurun_out/r2g_tests.txt
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo bench rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2g_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/r2g_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'dci_viterbi_kernel' -s 1 -c 1 -o gpurun_out/r2g_vit -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sub-records --pipelines 1 > gpurun_out/r2g_ncu_full.log 2>&1
cut -c1-400 gpurun_out/r2g_bench.json
