# how the round-2 profiles were produced (run with gpurun on one B200; outputs under gpurun_out/, summaries copied to profiles/r2_*)
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r2_tests.txt; tail -3 gpurun_out/r2_tests.txt
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo bench rc=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err; echo ref rc=$?

ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/r2_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'turbo_kernel|dci_viterbi_kernel|vit_worklist|ofdm_rx_kernel|chest_kernel|pdsch_demod_kernel|rm_turbo_rx|pdcch_llr_kernel|cand_compact_kernel|rb_power_kernel|scr_seq_kernel|tb_crc_kernel' -s 60 -c 22 -o gpurun_out/r2_all -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sub-records --pipelines 1 > gpurun_out/r2_ncu_full.log 2>&1
cut -c1-200 gpurun_out/r2_bench.json; cut -c1-200 gpurun_out/r2_bench_rmcb.json
