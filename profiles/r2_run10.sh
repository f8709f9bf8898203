# round 2, run 10: TMA-staged hop kernel with 1 KB rows (tile 2 x 4 sites x 256 columns); FP32-product Gram-Schmidt kernels
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "tma" 2>&1 | tail -5 > gpurun_out/r2_10_pytest_tma.log
tail -3 gpurun_out/r2_10_pytest_tma.log | cut -c1-300
for tma in 1 0; do
timeout 600 python bench.py --no-cpu-baseline --no-experiment --opt hop_tma=$tma > gpurun_out/r2_10_bench_tma_$tma.json 2> gpurun_out/r2_10_bench_tma_$tma.err || tail -5 gpurun_out/r2_10_bench_tma_$tma.err
python - <<PY
import json
d = json.load(open('gpurun_out/r2_10_bench_tma_$tma.json'))
r = d['roofline']
print('hop_tma=$tma', d['value'], d['e2e']['value'], d['fgmres_iters'], 'hop us', r['avg_launch_us'], 'frac', r['frac'], 'precond us', r['precondition_call_us'])
PY
done
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_10_pytest.log
tail -3 gpurun_out/r2_10_pytest.log | cut -c1-300
