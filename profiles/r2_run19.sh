# round 2, run 19 (1 GPU): GPU tier after the work-space fix and the two-column Gram-Schmidt kernels; bench with gs_x2 = 1 / 0
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2_19_pytest.log
grep -E "Error|passed|failed" gpurun_out/r2_19_pytest.log | head -8 | cut -c1-300
for x2 in 1 0; do
timeout 600 python bench.py --no-cpu-baseline --no-experiment --opt gs_x2=$x2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('gs_x2=$x2', d['value'], d['e2e']['value'], d['fgmres_iters'], d['gpu_launches'])"
done
