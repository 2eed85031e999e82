#!/bin/bash
# round-2 closing evidence on one B200 (run under gpurun from the repo root)
python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -6 > gpurun_out/r02y_pytest_gpu.log; cat gpurun_out/r02y_pytest_gpu.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02y_bench_reference_arm.json 2> gpurun_out/r02y_ref.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 3 > gpurun_out/r02y_bench_n1.json 2> gpurun_out/r02y_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02y_bench_n1.json') if l.startswith('{')][-1])
print('ms', d['ms_per_step'], 'clocks', d['clocks'])
e=d['e2e']; print('e2e', e['ms_per_step'], 'pageable', e['pageable']['ms_per_step'], 'long', e['long_call']['value'])
s=d['secondary']
for k,v in s.items(): print(k, json.dumps(v)[:300])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02y_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary > gpurun_out/r02y_ncu_launches.log 2>&1; echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:sgd_block -s 3 -c 1 -o gpurun_out/r02y_sgd python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary > gpurun_out/r02y_ncu_sgd.log 2>&1; echo "ncu sgd rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
