# round 2, run 27 (1 GPU): eigensolves on the set-up hierarchy itself, two-stage bootstrap on the coarse levels (their geometric
# hierarchies built early from the restricted fine vectors and kept): GPU tier, set-up profiles, synthetic 512^2 end to end
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2_27_pytest.log
grep -E "Error|passed|failed" gpurun_out/r2_27_pytest.log | head -8 | cut -c1-400
timeout 600 python profiles/profile_setup.py --L 512 --lines 40 > gpurun_out/r2_27_profile_setup_512.log 2>&1
grep -E "setup wall|test vectors level|Error|error" gpurun_out/r2_27_profile_setup_512.log | cut -c1-300
timeout 900 python profiles/profile_setup.py --L 1024 --lines 40 > gpurun_out/r2_27_profile_setup_1024.log 2>&1
grep -E "setup wall|test vectors level|Error|error" gpurun_out/r2_27_profile_setup_1024.log | cut -c1-300
timeout 900 python profiles/run_e2e.py --set synthetic512 --skip-hutchinson --batch 128 > gpurun_out/r2_27_synthetic512_1gpu.jsonl 2> gpurun_out/r2_27_synthetic512_1gpu.err
tail -3 gpurun_out/r2_27_synthetic512_1gpu.err | cut -c1-400
python - <<PY
import json
d = json.loads(open('gpurun_out/r2_27_synthetic512_1gpu.jsonl').read().strip().splitlines()[-1])
print('setup', d['setup_s'], 'sampling', d['sampling_s'], 'trace', d['trace'], [(l['nr_ests'], l['function_iters']) for l in d['levels']])
PY
