# round 2, run 26 (1 GPU): bench repeatability after the device set-up (two runs, JSON kept)
mkdir -p gpurun_out
for i in 1 2; do
timeout 600 python bench.py --no-cpu-baseline --no-experiment > gpurun_out/r2_26_bench_$i.json 2> gpurun_out/r2_26_bench_$i.err
python -c "
import json,sys
d=json.loads(open('gpurun_out/r2_26_bench_$i.json').read()); print(d['value'], d['e2e']['value'], d['fgmres_iters'], d['gpu_launches'], d.get('setup_s'), d['clocks'], d['roofline']['avg_launch_us'], d['roofline']['dense_umma_kernel']['us'])"
done
nvidia-smi --query-gpu=name,power.limit,clocks.max.sm,temperature.gpu --format=csv
