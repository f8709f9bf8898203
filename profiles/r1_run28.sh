mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r28_pytest.log
tail -4 gpurun_out/r28_pytest.log
python profiles/run_synthetic.py --L 256 --probes 128 > gpurun_out/r28_synth_256.json 2> gpurun_out/r28_synth_256.err; cat gpurun_out/r28_synth_256.json; tail -2 gpurun_out/r28_synth_256.err
python profiles/run_synthetic.py --L 512 --probes 64 > gpurun_out/r28_synth_512.json 2> gpurun_out/r28_synth_512.err; cat gpurun_out/r28_synth_512.json; tail -2 gpurun_out/r28_synth_512.err
python bench.py --steps 8 --warmup 3 > gpurun_out/r28_bench.json 2> gpurun_out/r28_bench.err; cut -c1-330 gpurun_out/r28_bench.json
