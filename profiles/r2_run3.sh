# round 2, run 3: device eigensolvers (test vectors, deflation), faster jump kernel on a low-priority stream, set-up profile
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2_3_pytest.log
tail -15 gpurun_out/r2_3_pytest.log
python -m pytest tests/test_eigensolve.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -30 > gpurun_out/r2_3_eigensolve.log
cat gpurun_out/r2_3_eigensolve.log
python profiles/time_mt_jump.py > gpurun_out/r2_3_mt_jump.jsonl 2> gpurun_out/r2_3_mt_jump.err || tail -5 gpurun_out/r2_3_mt_jump.err
cat gpurun_out/r2_3_mt_jump.jsonl
python profiles/profile_setup.py > gpurun_out/r2_3_profile_setup.log 2>&1; head -70 gpurun_out/r2_3_profile_setup.log
python bench.py --no-cpu-baseline > gpurun_out/r2_3_bench.json 2> gpurun_out/r2_3_bench.err || tail -20 gpurun_out/r2_3_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_3_bench.json'))
print(d['value'], d['e2e']['value'], d['e2e']['device_stream']['value'], d['experiment']['sampling_s'], d['experiment']['stop_indices'])
PY
