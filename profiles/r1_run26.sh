mkdir -p gpurun_out
python profiles/run_synthetic.py --L 256 --probes 128 > gpurun_out/r26_synth_256.json 2> gpurun_out/r26_synth_256.err; cat gpurun_out/r26_synth_256.json; tail -2 gpurun_out/r26_synth_256.err
python profiles/run_synthetic.py --L 512 --probes 64 > gpurun_out/r26_synth_512.json 2> gpurun_out/r26_synth_512.err; cat gpurun_out/r26_synth_512.json; tail -2 gpurun_out/r26_synth_512.err
