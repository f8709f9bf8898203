# round 2, run 6 (1 GPU): BASELINE configs[4] through the drivers (gateway.set_params('synthetic256') -> stoch_trace.mlmc), nothing
# injected (device eigensolvers), with the EXACT level traces (unit vectors) beside the estimate; then 512^2 on one GPU
mkdir -p gpurun_out
timeout 900 python profiles/run_e2e.py --set synthetic256 --skip-hutchinson --exact --batch 256 > gpurun_out/r2_6_synthetic256_1gpu.jsonl 2> gpurun_out/r2_6_synthetic256_1gpu.err
tail -3 gpurun_out/r2_6_synthetic256_1gpu.err; cut -c1-1500 gpurun_out/r2_6_synthetic256_1gpu.jsonl
timeout 900 python profiles/run_e2e.py --set synthetic512 --skip-hutchinson --batch 128 > gpurun_out/r2_6_synthetic512_1gpu.jsonl 2> gpurun_out/r2_6_synthetic512_1gpu.err
tail -3 gpurun_out/r2_6_synthetic512_1gpu.err; cut -c1-1500 gpurun_out/r2_6_synthetic512_1gpu.jsonl
