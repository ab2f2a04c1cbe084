set -x
cd $GRAFT_REPO_ROOT
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1d.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --pipelines 1 > gpurun_out/ncu_b1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:turbo_kernel -c 6 -f -o gpurun_out/prof_r1d_turbo python bench.py --steps 1 --warmup 3 --no-cpu-baseline --pipelines 1 > gpurun_out/ncu_b2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"dci_viterbi_kernel|cand_compact_kernel" -c 2 -f -o gpurun_out/prof_r1d_pdcch python bench.py --steps 1 --warmup 3 --no-cpu-baseline --pipelines 1 > gpurun_out/ncu_b3.log 2>&1
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo EXIT=$?
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo EXIT=$?
head -c 600 gpurun_out/bench_default.json; echo; head -c 600 gpurun_out/bench_reference.json
