# round 2, run 33 (2 GPUs): bench line at N = 2 with the final code (device set-up under torchrun)
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --no-cpu-baseline > gpurun_out/r2_33_bench_2gpu.json 2> gpurun_out/r2_33_bench_2gpu.err
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r2_33_bench_2gpu.json'))
    print(d['n_gpus'], d['value'], d['e2e']['value'], d['e2e']['device_stream']['value'], d['ms_per_step'], d['clocks'], d['experiment']['sampling_s'], d['experiment']['stop_indices'])
except Exception as e:
    print("bench 2gpu:", e)
PY
tail -3 gpurun_out/r2_33_bench_2gpu.err | cut -c1-300
