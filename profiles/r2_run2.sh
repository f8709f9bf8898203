# round 2, run 2: GPU tier with the MT19937 jump-ahead tests, default bench (outer_eo = 1 now the default; experiment G202,
# config-3 sweep, device-stream e2e, 16-probe CPU baseline), generator timings, ncu launch list + --set full at k = 512
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_2_pytest.log
tail -6 gpurun_out/r2_2_pytest.log
python bench.py > gpurun_out/r2_2_bench.json 2> gpurun_out/r2_2_bench.err || tail -20 gpurun_out/r2_2_bench.err
cut -c1-600 gpurun_out/r2_2_bench.json
python profiles/time_mt_jump.py > gpurun_out/r2_2_mt_jump.jsonl 2> gpurun_out/r2_2_mt_jump.err || tail -5 gpurun_out/r2_2_mt_jump.err
cat gpurun_out/r2_2_mt_jump.jsonl
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_2_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-experiment > gpurun_out/r2_2_ncu_bench.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"wilson_hop_eo_kernel|wilson_hop_eo_z|multi_dot_kernel|multi_axpy_norm|dense_umma|col_scale_eo" \
    --launch-skip 40 --launch-count 50 -o gpurun_out/r2_2_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-experiment > gpurun_out/r2_2_ncu_full.log 2>&1
ncu -i gpurun_out/r2_2_full.ncu-rep --page raw --csv > gpurun_out/r2_2_full_raw.csv 2>/dev/null
rm -f gpurun_out/r2_2_full.ncu-rep
ls -la gpurun_out | tail -12
