# round 2, run 17 (1 GPU): GPU tier (level-2 preconditioner test, slab-wise jump kernel in static shared memory), probe-stream
# overlap diagnostic (incl. high-priority variant), generator timings, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_17_pytest.log
tail -3 gpurun_out/r2_17_pytest.log | cut -c1-300
timeout 300 python profiles/time_mt_jump.py > gpurun_out/r2_17_mt_jump.jsonl 2>/dev/null; cat gpurun_out/r2_17_mt_jump.jsonl | grep jump-ahead
timeout 600 python profiles/probe_stream_overlap.py > gpurun_out/r2_17_probe_stream_overlap.json 2> gpurun_out/r2_17_probe_stream_overlap.err || tail -5 gpurun_out/r2_17_probe_stream_overlap.err
cat gpurun_out/r2_17_probe_stream_overlap.json
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_17_bench.json 2> gpurun_out/r2_17_bench.err || tail -5 gpurun_out/r2_17_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_17_bench.json'))
print(d['value'], d['e2e']['value'], d['e2e']['device_stream']['value'], d['experiment']['sampling_s'], d['experiment']['stop_indices'], d['setup_s'])
PY
