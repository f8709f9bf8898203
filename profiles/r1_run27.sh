mkdir -p gpurun_out
for cd in 36 16 8; do
python profiles/run_synthetic.py --L 256 --probes 128 --coarse-degree $cd > gpurun_out/r27_synth_256_cd$cd.json 2> gpurun_out/r27_synth_256_cd$cd.err; cat gpurun_out/r27_synth_256_cd$cd.json; tail -2 gpurun_out/r27_synth_256_cd$cd.err
done
