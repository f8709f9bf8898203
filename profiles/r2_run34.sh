# round 2, run 34 (1 GPU): last check of the committed tree -- GPU tier, smoke, the default bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_34_pytest.log
tail -1 gpurun_out/r2_34_pytest.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_34_smoke.log 2>&1; tail -1 gpurun_out/r2_34_smoke.log
timeout 600 python bench.py > gpurun_out/r2_34_bench.json 2> gpurun_out/r2_34_bench.err || tail -5 gpurun_out/r2_34_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_34_bench.json'))
print(d['value'], d['e2e']['value'], d['e2e']['device_stream']['value'], d['experiment']['sampling_s'], d['experiment']['stop_indices'], d['setup_s'], d['cpu_baseline']['value'], d['roofline']['frac'], d['roofline']['avg_launch_us'])
PY
grep -c "^\[build\]" gpurun_out/r2_34_bench.err
