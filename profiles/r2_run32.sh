# round 2, run 32 (1 GPU): synthetic 512^2 through the drivers on ONE GPU with the code of run 31 (same digits as on eight?)
mkdir -p gpurun_out
timeout 600 python profiles/run_e2e.py --set synthetic512 --skip-hutchinson --batch 128 > gpurun_out/r2_32_synthetic512_1gpu.jsonl 2> gpurun_out/r2_32_synthetic512_1gpu.err
tail -2 gpurun_out/r2_32_synthetic512_1gpu.err | cut -c1-300; tail -1 gpurun_out/r2_32_synthetic512_1gpu.jsonl | cut -c1-1500
