python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2f_tests.txt; tail -4 gpurun_out/r2f_tests.txt
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo bench rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/r2f_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'dci_viterbi_kernel|ofdm_rx_kernel|pusch_kernel' -s 2 -c 4 -o gpurun_out/r2f_vit_ofdm -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sub-records --pipelines 1 > gpurun_out/r2f_ncu_full.log 2>&1
tail -c 400 gpurun_out/r2f_bench.err; cut -c1-600 gpurun_out/r2f_bench.json
