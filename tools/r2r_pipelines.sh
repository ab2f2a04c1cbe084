for T in 4 6 8; do
  python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-sub-records --pipelines $T > gpurun_out/r2r_T$T.json 2> gpurun_out/r2r_T$T.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2r_T$T.json').read().strip().splitlines()[-1])
print('T=$T value %.0f (%.2f ms)  e2e %.0f (%.2f ms)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']))
PY
done
