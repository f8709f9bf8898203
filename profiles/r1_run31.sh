mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r31_pytest.log
tail -4 gpurun_out/r31_pytest.log
python profiles/tune_geometric.py '[{"degree": 36}, {"degree": 36, "fuse_res": 0}, {"degree": 36, "adaptive_poll": 0}, {"degree": 34}, {"degree": 40}]' > gpurun_out/r31_tune.jsonl 2> gpurun_out/r31_tune.err
cat gpurun_out/r31_tune.jsonl; tail -3 gpurun_out/r31_tune.err
