# round 2, run 31 (8 GPUs): BASELINE configs[4] through the drivers with the device set-up (two-stage eigensolves, device Galerkin
# products): synthetic 1024^2 and 512^2 deflated MLMC to the variance target, nothing injected, probes sharded x8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29513 profiles/run_e2e.py --set synthetic1024 --skip-hutchinson --batch 32 > gpurun_out/r2_31_synthetic1024_8gpu.jsonl 2> gpurun_out/r2_31_synthetic1024_8gpu.err
tail -2 gpurun_out/r2_31_synthetic1024_8gpu.err | cut -c1-300; tail -1 gpurun_out/r2_31_synthetic1024_8gpu.jsonl | cut -c1-1500
timeout 400 $TR --master-port 29512 profiles/run_e2e.py --set synthetic512 --skip-hutchinson --batch 128 > gpurun_out/r2_31_synthetic512_8gpu.jsonl 2> gpurun_out/r2_31_synthetic512_8gpu.err
tail -2 gpurun_out/r2_31_synthetic512_8gpu.err | cut -c1-300; tail -1 gpurun_out/r2_31_synthetic512_8gpu.jsonl | cut -c1-1500
