# round 2, run 16 (8 GPUs): BASELINE configs[4] through the drivers with the level-2 geometric preconditioner hierarchy:
# synthetic 512^2 and 1024^2 deflated MLMC to the variance target, nothing injected, probes sharded x8; then the bench line
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29512 profiles/run_e2e.py --set synthetic512 --skip-hutchinson --batch 128 > gpurun_out/r2_16_synthetic512_8gpu.jsonl 2> gpurun_out/r2_16_synthetic512_8gpu.err
tail -2 gpurun_out/r2_16_synthetic512_8gpu.err | cut -c1-300; tail -1 gpurun_out/r2_16_synthetic512_8gpu.jsonl | cut -c1-1200
timeout 900 $TR --master-port 29513 profiles/run_e2e.py --set synthetic1024 --skip-hutchinson --batch 32 > gpurun_out/r2_16_synthetic1024_8gpu.jsonl 2> gpurun_out/r2_16_synthetic1024_8gpu.err
tail -2 gpurun_out/r2_16_synthetic1024_8gpu.err | cut -c1-300; tail -1 gpurun_out/r2_16_synthetic1024_8gpu.jsonl | cut -c1-1200
timeout 600 $TR --master-port 29511 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/r2_16_bench_8gpu.json 2> gpurun_out/r2_16_bench_8gpu.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2_16_bench_8gpu.json'))
    print(d['n_gpus'], d['value'], d['e2e']['value'], d['e2e']['device_stream']['value'], d['ms_per_step'], d['clocks'], d['experiment']['sampling_s'], d['experiment']['stop_indices'])
except Exception as e:
    print("bench 8gpu:", e)
PY
