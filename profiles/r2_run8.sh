# round 2, run 8: outer_drop sweep of the mixed-precision solve, device prolongator values, launch list + ncu of the new default
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2_8_pytest.log
tail -5 gpurun_out/r2_8_pytest.log
for drop in 1e-3 3e-4 1e-4 3e-5; do
python bench.py --no-cpu-baseline --no-experiment --opt outer_drop=$drop 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('outer_drop=$drop', d['value'], d['e2e']['value'], d['fgmres_iters'], d['gpu_launches'])" | tee -a gpurun_out/r2_8_outer_drop_sweep.txt
done
python bench.py > gpurun_out/r2_8_bench.json 2> gpurun_out/r2_8_bench.err || tail -20 gpurun_out/r2_8_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_8_bench.json'))
print(d['value'], d['e2e']['value'], d['e2e']['device_stream']['value'], d['experiment']['sampling_s'], d['experiment']['stop_indices'], d['setup_s'], d['roofline']['frac'])
PY
python profiles/profile_setup.py --lines 25 > gpurun_out/r2_8_profile_setup_128.log 2>&1; head -32 gpurun_out/r2_8_profile_setup_128.log | cut -c1-160
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_8_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-experiment > gpurun_out/r2_8_ncu_bench.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"wilson_hop_eo_kernel|wilson_hop_eo_z|multi_dot_kernel|multi_axpy_norm|dense_umma|col_scale_eo|stencil_kernel<float" \
    --launch-skip 40 --launch-count 60 -o gpurun_out/r2_8_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-experiment > gpurun_out/r2_8_ncu_full.log 2>&1
ncu -i gpurun_out/r2_8_full.ncu-rep --page raw --csv > gpurun_out/r2_8_full_raw.csv 2>/dev/null
rm -f gpurun_out/r2_8_full.ncu-rep
