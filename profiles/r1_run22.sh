# run 22: bench + ncu launch list + ncu full captures of the geometric-preconditioner configuration (degree 32)
set -x
mkdir -p gpurun_out
python bench.py --steps 8 --warmup 3 > gpurun_out/r22_bench.json 2> gpurun_out/r22_bench.err || exit 1
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r22_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r22_ncu_bench.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"stencil_step_bf16_t2|multi_dot_kernel|multi_axpy_norm" \
    --launch-skip 40 --launch-count 6 -o gpurun_out/r22_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r22_ncu_full.log 2>&1
ncu -i gpurun_out/r22_full.ncu-rep --page raw --csv > gpurun_out/r22_full_raw.csv 2>/dev/null
ls -la gpurun_out | tail -8
cut -c1-300 gpurun_out/r22_bench.json
