# round 2, run 29 (1 GPU): final state -- GPU tier, smoke, full bench line, launch list and ncu --set full of the timed region,
# ncu --set full of the config-3 operators (SpMM on every level, restriction, prolongation) at k = 256 / 512
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2_29_pytest.log
tail -3 gpurun_out/r2_29_pytest.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_29_smoke.log 2>&1; tail -1 gpurun_out/r2_29_smoke.log
timeout 900 python bench.py > gpurun_out/r2_29_bench.json 2> gpurun_out/r2_29_bench.err || tail -5 gpurun_out/r2_29_bench.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/r2_29_bench.json'))
print(d['value'], d['e2e']['value'], d['e2e']['device_stream']['value'], d['experiment']['sampling_s'], d['experiment']['stop_indices'], d['setup_s'], d['cpu_baseline']['value'], d['roofline']['frac'], d['roofline']['avg_launch_us'])
PY
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_29_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-experiment > gpurun_out/r2_29_ncu_bench.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"wilson_hop_eo_kernel|wilson_hop_eo_z|wilson_schur_residual|multi_dot|multi_axpy_norm|dense_umma|col_scale_eo|stencil_kernel<float|prolong_add_kernel<float|restrict_kernel<float" \
    --launch-skip 40 --launch-count 70 -o gpurun_out/r2_29_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-experiment > gpurun_out/r2_29_ncu_full.log 2>&1
ncu -i gpurun_out/r2_29_full.ncu-rep --page raw --csv > gpurun_out/r2_29_full_raw.csv 2>/dev/null
rm -f gpurun_out/r2_29_full.ncu-rep
ncu --profile-from-start off --set full --clock-control none \
    -k regex:"stencil_kernel|bsr_kernel|restrict_kernel|prolong_add_kernel" \
    -o gpurun_out/r2_29_spmm python profiles/spmm_ncu.py > gpurun_out/r2_29_ncu_spmm.log 2>&1
ncu -i gpurun_out/r2_29_spmm.ncu-rep --page raw --csv > gpurun_out/r2_29_spmm_raw.csv 2>/dev/null
rm -f gpurun_out/r2_29_spmm.ncu-rep
ls -la gpurun_out/r2_29_*
