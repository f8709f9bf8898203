# round 2, run 5: GPU tier; un-injected set-up (device eigensolver on every level) at 128^2 and on synthetic 512^2 / 1024^2
# with the set-up profile; bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2_5_pytest.log
tail -5 gpurun_out/r2_5_pytest.log
python profiles/profile_setup.py --uninjected --lines 30 > gpurun_out/r2_5_profile_setup_128_uninjected.log 2>&1; head -50 gpurun_out/r2_5_profile_setup_128_uninjected.log
timeout 900 python profiles/profile_setup.py --L 512 --lines 40 > gpurun_out/r2_5_profile_setup_512.log 2>&1; head -60 gpurun_out/r2_5_profile_setup_512.log
python bench.py --no-cpu-baseline > gpurun_out/r2_5_bench.json 2> gpurun_out/r2_5_bench.err || tail -20 gpurun_out/r2_5_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_5_bench.json'))
print(d['value'], d['e2e']['value'], d['e2e']['device_stream']['value'], d['experiment']['sampling_s'], d['experiment']['stop_indices'], d['setup_s'])
PY
