# round 2, run 24 (1 GPU): two-stage bootstrap test at 128^2, GPU tier, set-up profiles 512^2 / 1024^2 (nothing injected)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_eigensolve.py -m gpu -x -q -s 2>&1 | tail -15 | cut -c1-400 > gpurun_out/r2_24_pytest_eigensolve.log
grep -E "two-stage|Error|passed|failed" gpurun_out/r2_24_pytest_eigensolve.log | head -8
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2_24_pytest.log
grep -E "Error|passed|failed" gpurun_out/r2_24_pytest.log | head -8 | cut -c1-400
timeout 600 python profiles/profile_setup.py --L 512 --lines 40 > gpurun_out/r2_24_profile_setup_512.log 2>&1
grep -E "setup wall|test vectors level|Error|error" gpurun_out/r2_24_profile_setup_512.log | cut -c1-300
timeout 900 python profiles/profile_setup.py --L 1024 --lines 40 > gpurun_out/r2_24_profile_setup_1024.log 2>&1
grep -E "setup wall|test vectors level|Error|error" gpurun_out/r2_24_profile_setup_1024.log | cut -c1-300
