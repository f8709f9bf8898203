# round 2, run 4: full GPU tier after the set-up changes (device Arnoldi / storage check / inverses), set-up profile, bench with
# the 32-chunk generator, reference arm on the unmodified reference (oracle/_ref)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2_4_pytest.log
tail -8 gpurun_out/r2_4_pytest.log
python -m pytest tests/test_eigensolve.py -m gpu -q -s 2>&1 | grep -v "^$" | tail -12 > gpurun_out/r2_4_eigensolve.log
cat gpurun_out/r2_4_eigensolve.log
python profiles/profile_setup.py > gpurun_out/r2_4_profile_setup.log 2>&1; head -40 gpurun_out/r2_4_profile_setup.log
python bench.py > gpurun_out/r2_4_bench.json 2> gpurun_out/r2_4_bench.err || tail -20 gpurun_out/r2_4_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_4_bench.json'))
print(d['value'], d['e2e']['value'], d['e2e']['device_stream']['value'], d['experiment']['sampling_s'], d['experiment']['stop_indices'], d['setup_s'])
print(d['cpu_baseline'])
PY
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_4_bench_reference_arm.json 2> gpurun_out/r2_4_bench_reference_arm.err
cut -c1-300 gpurun_out/r2_4_bench_reference_arm.json; tail -3 gpurun_out/r2_4_bench_reference_arm.err
