python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo bench rc=$?
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2d_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sub-records > gpurun_out/r2d_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'turbo_kernel|dci_viterbi_kernel' -s 14 -c 10 -o gpurun_out/r2d_turbo_vit -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sub-records --pipelines 1 > gpurun_out/r2d_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'ofdm|chest|pdsch_demod|rate_dematch|pdcch|cand_compact|pack' -s 20 -c 12 -o gpurun_out/r2d_others -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-sub-records --pipelines 1 > gpurun_out/r2d_ncu_full2.log 2>&1
ls -la gpurun_out/ | tail -12; tail -c 600 gpurun_out/r2d_bench.err
