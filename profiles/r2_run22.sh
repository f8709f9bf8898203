# round 2, run 22 (1 GPU): Gauss-Jordan as two kernels per pivot (set-up tests, set-up profile); two-stream overlap experiment
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_setup.py -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2_22_pytest_setup.log
grep -E "Error|passed|failed" gpurun_out/r2_22_pytest_setup.log | head -8 | cut -c1-400
timeout 300 python profiles/profile_setup.py --lines 30 > gpurun_out/r2_22_profile_setup_128.log 2>&1
grep "setup wall" gpurun_out/r2_22_profile_setup_128.log
timeout 600 python profiles/exp_two_streams.py --k 256 > gpurun_out/r2_22_two_streams_256.json 2> gpurun_out/r2_22_two_streams.err
cat gpurun_out/r2_22_two_streams_256.json; tail -3 gpurun_out/r2_22_two_streams.err
timeout 600 python profiles/exp_two_streams.py --k 512 > gpurun_out/r2_22_two_streams_512.json 2>> gpurun_out/r2_22_two_streams.err
cat gpurun_out/r2_22_two_streams_512.json
