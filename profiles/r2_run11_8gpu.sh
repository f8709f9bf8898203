# round 2, run 11 (8 GPUs): bench line with the strong-scaling experiment G202, then BASELINE configs[4] through the drivers:
# synthetic 512^2 and 1024^2 deflated MLMC to the variance target, nothing injected (device eigensolvers), probes sharded x8
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/r2_11_bench_8gpu.json 2> gpurun_out/r2_11_bench_8gpu.err
tail -2 gpurun_out/r2_11_bench_8gpu.err | cut -c1-300
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2_11_bench_8gpu.json'))
    print(d['n_gpus'], d['value'], d['e2e']['value'], d['e2e']['device_stream']['value'], d['experiment'])
except Exception as e:
    print("bench 8gpu:", e)
PY
timeout 600 $TR --master-port 29512 profiles/run_e2e.py --set synthetic512 --skip-hutchinson --batch 128 > gpurun_out/r2_11_synthetic512_8gpu.jsonl 2> gpurun_out/r2_11_synthetic512_8gpu.err
tail -2 gpurun_out/r2_11_synthetic512_8gpu.err | cut -c1-300; cut -c1-1200 gpurun_out/r2_11_synthetic512_8gpu.jsonl
timeout 900 $TR --master-port 29513 profiles/run_e2e.py --set synthetic1024 --skip-hutchinson --batch 32 > gpurun_out/r2_11_synthetic1024_8gpu.jsonl 2> gpurun_out/r2_11_synthetic1024_8gpu.err
tail -2 gpurun_out/r2_11_synthetic1024_8gpu.err | cut -c1-300; cut -c1-1200 gpurun_out/r2_11_synthetic1024_8gpu.jsonl
