# round 2, run 12: full GPU tier on the final code, smoke, batch-size sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_12_pytest.log
tail -3 gpurun_out/r2_12_pytest.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_12_smoke.log 2>&1; tail -2 gpurun_out/r2_12_smoke.log
for k in 256 768 1024; do
timeout 600 python bench.py --no-cpu-baseline --no-experiment --probes $k > gpurun_out/r2_12_bench_k$k.json 2> gpurun_out/r2_12_bench_k$k.err || tail -5 gpurun_out/r2_12_bench_k$k.err
python - <<PY
import json
d = json.load(open('gpurun_out/r2_12_bench_k$k.json'))
r = d['roofline']
print('k=$k', d['value'], d['e2e']['value'], d['e2e']['device_stream']['value'], d['fgmres_iters'], 'hop us', r['avg_launch_us'], 'frac', r['frac'], 'clock query ms', d['clocks']['query_ms_max'])
PY
done
