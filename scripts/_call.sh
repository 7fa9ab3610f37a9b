set -x
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-other-configs > gpurun_out/bench_arena.log 2>&1; python - <<'PY'
import json
d=json.loads([x for x in open('gpurun_out/bench_arena.log') if x.startswith('{')][-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'])
PY
python -m pytest tests -m gpu -q -x -k "frames or golden or edge or ragged or bad or upload or multi or example or slices or program" 2>&1 | tail -3
