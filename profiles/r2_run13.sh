# round 2, run 13 (1 GPU): geometric preconditioner hierarchy for the estimator's level-2 solves (synthetic 512^2), on / off
mkdir -p gpurun_out
for cg in 2 1; do
timeout 900 python profiles/run_e2e.py --set synthetic512 --skip-hutchinson --batch 128 --coarse-geo $cg > gpurun_out/r2_13_synthetic512_coarse_geo_$cg.jsonl 2> gpurun_out/r2_13_synthetic512_coarse_geo_$cg.err
tail -3 gpurun_out/r2_13_synthetic512_coarse_geo_$cg.err | cut -c1-400
python - <<PY
import json
d = json.loads(open('gpurun_out/r2_13_synthetic512_coarse_geo_$cg.jsonl').read().strip().splitlines()[-1])
print('coarse_geo=$cg setup', d['setup_s'], 'sampling', d['sampling_s'], 'trace', d['trace'], [(l['nr_ests'], l['function_iters']) for l in d['levels']])
PY
done
