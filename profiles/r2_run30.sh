# round 2, run 30 (1 GPU): row chunk of the complex64 Gram-Schmidt kernels (automatic / 256 / 128 / 64 / 32): GPU tier, bench sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_30_pytest.log
tail -3 gpurun_out/r2_30_pytest.log | cut -c1-300
for rows in 0 256 128 64 32; do
timeout 600 python bench.py --no-cpu-baseline --no-experiment --opt gs_rows=$rows 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('gs_rows=$rows', d['value'], d['e2e']['value'], d['fgmres_iters'], d['gpu_launches'])" | tee -a gpurun_out/r2_30_gs_rows_sweep.txt
done
