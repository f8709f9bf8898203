# round 2, run 25 (1 GPU): even-odd smoother set-up on the device, start blocks drawn on the device: tests, set-up profiles
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_setup.py tests/test_eigensolve.py -m gpu -x -q 2>&1 | tail -15 | cut -c1-400 > gpurun_out/r2_25_pytest_setup.log
grep -E "Error|passed|failed" gpurun_out/r2_25_pytest_setup.log | head -8
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2_25_pytest.log
grep -E "Error|passed|failed" gpurun_out/r2_25_pytest.log | head -8 | cut -c1-400
timeout 300 python profiles/profile_setup.py --lines 30 > gpurun_out/r2_25_profile_setup_128.log 2>&1
grep "setup wall" gpurun_out/r2_25_profile_setup_128.log
timeout 600 python profiles/profile_setup.py --L 512 --lines 40 > gpurun_out/r2_25_profile_setup_512.log 2>&1
grep -E "setup wall|Error|error" gpurun_out/r2_25_profile_setup_512.log | cut -c1-300
timeout 900 python profiles/profile_setup.py --L 1024 --lines 40 > gpurun_out/r2_25_profile_setup_1024.log 2>&1
grep -E "setup wall|Error|error" gpurun_out/r2_25_profile_setup_1024.log | cut -c1-300
timeout 600 python bench.py --no-cpu-baseline --no-experiment 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['fgmres_iters'], d['gpu_launches'], d.get('setup_s'))"
