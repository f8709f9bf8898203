# run 33: final configuration of the round: tests, bench, reference arm, ncu launch list, ncu full of the top kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r33_pytest.log
tail -4 gpurun_out/r33_pytest.log
python bench.py --steps 8 --warmup 3 > gpurun_out/r33_bench.json 2> gpurun_out/r33_bench.err || tail -5 gpurun_out/r33_bench.err
cut -c1-330 gpurun_out/r33_bench.json
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r33_smoke.log 2>&1; tail -2 gpurun_out/r33_smoke.log
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r33_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r33_ncu_bench.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"multi_dot_kernel|multi_axpy_norm|dense_umma" \
    --launch-skip 6 --launch-count 5 -o gpurun_out/r33_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r33_ncu_full.log 2>&1
ncu -i gpurun_out/r33_full.ncu-rep --page raw --csv > gpurun_out/r33_full_raw.csv 2>/dev/null
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r33_bench_reference.json 2> gpurun_out/r33_bench_reference.err; cut -c1-300 gpurun_out/r33_bench_reference.json
