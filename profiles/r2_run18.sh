# round 2, run 18 (1 GPU): final state -- GPU tier, smoke, bench (full), launch list of the timed region
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_18_pytest.log
tail -3 gpurun_out/r2_18_pytest.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_18_smoke.log 2>&1; tail -1 gpurun_out/r2_18_smoke.log
timeout 900 python bench.py > gpurun_out/r2_18_bench.json 2> gpurun_out/r2_18_bench.err || tail -5 gpurun_out/r2_18_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_18_bench.json'))
print(d['value'], d['e2e']['value'], d['e2e']['device_stream']['value'], d['experiment']['sampling_s'], d['experiment']['stop_indices'], d['setup_s'], d['cpu_baseline']['value'], d['roofline']['frac'], d['roofline']['avg_launch_us'])
PY
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_18_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-experiment > gpurun_out/r2_18_ncu_bench.log 2>&1
