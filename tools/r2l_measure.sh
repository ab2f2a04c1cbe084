python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2l_tests.txt; tail -4 gpurun_out/r2l_tests.txt
python bench.py > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo bench rc=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2l_bench_ref.json 2> gpurun_out/r2l_bench_ref.err; echo ref rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2l_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/r2l_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'turbo_kernel|dci_viterbi_kernel|ofdm_rx_kernel|chest_kernel|pdsch_demod_kernel|rm_turbo_rx_kernel|pdcch_llr_kernel|cand_compact_kernel|rb_power_kernel|scr_seq_kernel|tb_crc_kernel|pull_kernel' -s 40 -c 24 -o gpurun_out/r2l_all -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sub-records --pipelines 1 > gpurun_out/r2l_ncu_full.log 2>&1
cut -c1-300 gpurun_out/r2l_bench.json
