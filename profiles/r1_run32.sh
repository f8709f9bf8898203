mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r32_pytest.log
tail -4 gpurun_out/r32_pytest.log
python profiles/tune_geometric.py '[{"degree": 36}, {"degree": 36, "dot32": 0}, {"degree": 36}]' > gpurun_out/r32_tune.jsonl 2> gpurun_out/r32_tune.err
cat gpurun_out/r32_tune.jsonl; tail -3 gpurun_out/r32_tune.err
