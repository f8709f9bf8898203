# round 2, run 15 (1 GPU): where the device-stream e2e loses 10 % (profiles/probe_stream_overlap.py)
mkdir -p gpurun_out
timeout 600 python profiles/probe_stream_overlap.py > gpurun_out/r2_15_probe_stream_overlap.json 2> gpurun_out/r2_15_probe_stream_overlap.err || tail -5 gpurun_out/r2_15_probe_stream_overlap.err
cat gpurun_out/r2_15_probe_stream_overlap.json
