set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_frames_in_flight.py -q -m gpu > gpurun_out/r04b_fif_tests.log 2>&1; tail -15 gpurun_out/r04b_fif_tests.log
timeout 600 python bench.py > gpurun_out/r04b_bench_default.json 2> gpurun_out/r04b_bench_default.err; tail -3 gpurun_out/r04b_bench_default.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r04b_bench_default.json'))
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'fif', d['frames_in_flight'], 'frac', d['roofline']['frac'], d['clocks'], 'launches', d['gpu_launches'])
for k,v in d['extra']['configs'].items(): print(k, round(v['value'],1), round(v['ms_per_step'],2), 'e2e', round(v['e2e']['value'],1), v['frames_in_flight'])
PY
timeout 300 python bench.py --no-extra --no-cpu-baseline --frames-in-flight 1 > gpurun_out/r04b_bench_fif1.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r04b_bench_fif1.json')); print('fif1', d['value'], d['e2e']['value'])"
