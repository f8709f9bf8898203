# round 2, run 21 (1 GPU): set-up on the device (Galerkin product, geometric Gram-Schmidt, coarsest inverse): new tests, GPU tier,
# set-up profile at 128^2 (injected) and bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_setup.py -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2_21_pytest_setup.log
grep -E "Error|passed|failed" gpurun_out/r2_21_pytest_setup.log | head -8 | cut -c1-400
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2_21_pytest.log
grep -E "Error|passed|failed" gpurun_out/r2_21_pytest.log | head -8 | cut -c1-400
timeout 300 python profiles/profile_setup.py --lines 30 > gpurun_out/r2_21_profile_setup_128.log 2>&1
grep "setup wall" gpurun_out/r2_21_profile_setup_128.log
timeout 600 python bench.py --no-cpu-baseline --no-experiment 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['fgmres_iters'], d['gpu_launches'], d.get('setup_s'))"
