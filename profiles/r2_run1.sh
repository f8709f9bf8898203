# round 2, run 1: full GPU tier (new parity tests, Schur outer solve first time on hardware), bench at the default k = 512
# with and without outer_eo, ncu launch list + ncu --set full of the five top kernels at k = 512
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_1_pytest.log
tail -6 gpurun_out/r2_1_pytest.log
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r2_1_bench.json 2> gpurun_out/r2_1_bench.err || tail -5 gpurun_out/r2_1_bench.err
cut -c1-400 gpurun_out/r2_1_bench.json
python bench.py --steps 8 --warmup 3 --no-cpu-baseline --opt outer_eo=1 > gpurun_out/r2_1_bench_outer_eo.json 2> gpurun_out/r2_1_bench_outer_eo.err || tail -5 gpurun_out/r2_1_bench_outer_eo.err
cut -c1-400 gpurun_out/r2_1_bench_outer_eo.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_1_smoke.log 2>&1; tail -2 gpurun_out/r2_1_smoke.log
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_1_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_1_ncu_bench.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on \
    -k regex:"wilson_hop_eo_kernel|stencil_kernel<double|multi_dot_kernel|multi_axpy_norm|dense_umma" \
    --launch-skip 40 --launch-count 45 -o gpurun_out/r2_1_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_1_ncu_full.log 2>&1
ncu -i gpurun_out/r2_1_full.ncu-rep --page raw --csv > gpurun_out/r2_1_full_raw.csv 2>/dev/null
ls -la gpurun_out | tail -12
