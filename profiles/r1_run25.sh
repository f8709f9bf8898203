mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r25_pytest.log
tail -4 gpurun_out/r25_pytest.log
python profiles/tune_geometric.py > gpurun_out/r25_tune_geometric.jsonl 2> gpurun_out/r25_tune.err
cat gpurun_out/r25_tune_geometric.jsonl; tail -3 gpurun_out/r25_tune.err
