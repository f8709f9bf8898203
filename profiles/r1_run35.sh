mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r35_pytest.log
tail -5 gpurun_out/r35_pytest.log
python profiles/tune_geometric.py '[{"degree": 36, "eo_degree": 16}, {"degree": 36, "eo_degree": 16, "eo_packs": 1}, {"degree": 36, "eo_degree": 16, "eo_by": 2, "eo_bz": 4}, {"degree": 36, "eo_degree": 16, "eo_by": 8, "eo_bz": 1}, {"degree": 36, "eo_degree": 16, "eo_by": 2, "eo_bz": 2}, {"degree": 36, "eo_degree": 18}]' > gpurun_out/r35_tune.jsonl 2> gpurun_out/r35_tune.err
cat gpurun_out/r35_tune.jsonl; tail -3 gpurun_out/r35_tune.err
