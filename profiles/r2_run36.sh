# round 2, run 36 (1 GPU): GPU tier of the final tree
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_36_pytest.log
tail -1 gpurun_out/r2_36_pytest.log | cut -c1-300
