# round 2, run 35 (1 GPU): synthetic 1024^2, does a geometric preconditioner hierarchy for the level-3 solves (n = 32 768, coarsest
# level 8 192 inverted on the device) pay?  geometric_coarse_levels = 2 (default) against 3, trace_tol 3e-2 to keep the runs short
mkdir -p gpurun_out
for cg in 2 3; do
timeout 170 python profiles/run_e2e.py --set synthetic1024 --skip-hutchinson --batch 32 --trace-tol 3e-2 --coarse-geo $cg > gpurun_out/r2_35_synthetic1024_coarse_geo_$cg.jsonl 2> gpurun_out/r2_35_synthetic1024_coarse_geo_$cg.err
tail -2 gpurun_out/r2_35_synthetic1024_coarse_geo_$cg.err | cut -c1-300
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/r2_35_synthetic1024_coarse_geo_$cg.jsonl').read().strip().splitlines()[-1])
    print('coarse_geo=$cg setup', d['setup_s'], 'sampling', d['sampling_s'], 'trace', d['trace'], [(l['nr_ests'], l['function_iters']) for l in d['levels']])
except Exception as e:
    print('coarse_geo=$cg failed:', e)
PY
done
