# round 2, run 23 (1 GPU): two-stage bootstrap of the level-0 test-vector eigensolve (synthetic 512^2 set-up profile, on / off),
# then the whole synthetic 512^2 experiment through the drivers
mkdir -p gpurun_out
timeout 600 python profiles/profile_setup.py --L 512 --lines 25 > gpurun_out/r2_23_profile_setup_512_two_stage.log 2>&1
grep -E "setup wall|test vectors level|Error|error" gpurun_out/r2_23_profile_setup_512_two_stage.log | cut -c1-300
timeout 600 python profiles/profile_setup.py --L 512 --lines 25 --no-two-stage > gpurun_out/r2_23_profile_setup_512_one_stage.log 2>&1
grep -E "setup wall|test vectors level|Error|error" gpurun_out/r2_23_profile_setup_512_one_stage.log | cut -c1-300
timeout 900 python profiles/run_e2e.py --set synthetic512 --skip-hutchinson --batch 128 > gpurun_out/r2_23_synthetic512_1gpu.jsonl 2> gpurun_out/r2_23_synthetic512_1gpu.err
tail -3 gpurun_out/r2_23_synthetic512_1gpu.err | cut -c1-400
python - <<PY
import json
d = json.loads(open('gpurun_out/r2_23_synthetic512_1gpu.jsonl').read().strip().splitlines()[-1])
print('setup', d['setup_s'], 'sampling', d['sampling_s'], 'trace', d['trace'], [(l['nr_ests'], l['function_iters']) for l in d['levels']])
PY
