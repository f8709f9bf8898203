# round 2, run 20 (1 GPU): fused Schur residual + norm kernel, two-column normalisation kernel: GPU tier, bench on / off
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2_20_pytest.log
grep -E "Error|passed|failed" gpurun_out/r2_20_pytest.log | head -8 | cut -c1-300
for opt in "fuse_residual=1" "fuse_residual=0" "gs_x2=0"; do
timeout 600 python bench.py --no-cpu-baseline --no-experiment --opt $opt 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$opt', d['value'], d['e2e']['value'], d['fgmres_iters'], d['gpu_launches'])"
done
