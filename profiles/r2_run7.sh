# round 2, run 7: mixed-precision Schur-complement solve (Krylov vectors stored in complex64, option outer_c64, default 1):
# GPU tier, bench with outer_c64 = 1 / 0, cycle structure
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2_7_pytest.log
tail -8 gpurun_out/r2_7_pytest.log
for c64 in 1 0; do
python bench.py --no-cpu-baseline --no-experiment --opt outer_c64=$c64 > gpurun_out/r2_7_bench_c64_$c64.json 2> gpurun_out/r2_7_bench_c64_$c64.err || tail -20 gpurun_out/r2_7_bench_c64_$c64.err
python - <<PY
import json
d = json.load(open('gpurun_out/r2_7_bench_c64_$c64.json'))
print('outer_c64=$c64', d['value'], d['e2e']['value'], d['e2e']['device_stream']['value'], d['fgmres_iters'], d['gpu_launches'])
PY
done
for drop in 1e-4 1e-6; do
python bench.py --no-cpu-baseline --no-experiment --opt outer_drop=$drop 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('outer_drop=$drop', d['value'], d['fgmres_iters'], d['gpu_launches'])"
done
