mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r34_pytest.log
tail -5 gpurun_out/r34_pytest.log
python profiles/tune_geometric.py '[{"degree": 36, "eo_degree": 16}, {"degree": 36, "smoother_eo": 0}, {"degree": 36, "eo_degree": 12}, {"degree": 36, "eo_degree": 14}, {"degree": 36, "eo_degree": 18}, {"degree": 36, "eo_degree": 20}, {"degree": 36, "eo_degree": 16, "eo_by": 2, "eo_bz": 4}, {"degree": 36, "eo_degree": 16, "eo_by": 8, "eo_bz": 1}]' > gpurun_out/r34_tune.jsonl 2> gpurun_out/r34_tune.err
cat gpurun_out/r34_tune.jsonl; tail -3 gpurun_out/r34_tune.err
