# run 37: bench at the default batch (512 probes per step), batch 768 / 1024, launch list at 512
mkdir -p gpurun_out
python bench.py --steps 8 --warmup 3 > gpurun_out/r37_bench.json 2> gpurun_out/r37_bench.err || tail -5 gpurun_out/r37_bench.err
cut -c1-330 gpurun_out/r37_bench.json
for k in 768 1024; do
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --probes $k > gpurun_out/r37_bench_k$k.json 2> gpurun_out/r37_bench_k$k.err; cut -c1-260 gpurun_out/r37_bench_k$k.json; tail -2 gpurun_out/r37_bench_k$k.err
done
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r37_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r37_ncu_bench.log 2>&1
