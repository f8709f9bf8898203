# round 2, run 14 (1 GPU): GPU tier after the coarse-level preconditioners and the iteration-count bookkeeping; synthetic 256^2
# with the exact traces again (the parity anchor of configs[4]); bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_14_pytest.log
tail -3 gpurun_out/r2_14_pytest.log | cut -c1-300
timeout 600 python bench.py > gpurun_out/r2_14_bench.json 2> gpurun_out/r2_14_bench.err || tail -5 gpurun_out/r2_14_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_14_bench.json'))
print(d['value'], d['e2e']['value'], d['e2e']['device_stream']['value'], d['experiment']['sampling_s'], d['experiment']['stop_indices'], d['setup_s'], d['cpu_baseline']['value'], d['cpu_baseline']['kind'])
PY
