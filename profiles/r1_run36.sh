# run 36: bench line with the roofline of the new dominant kernel, ncu --set full of wilson_hop_eo_kernel, batch 512 / 384
mkdir -p gpurun_out
python bench.py --steps 8 --warmup 3 > gpurun_out/r36_bench.json 2> gpurun_out/r36_bench.err || tail -5 gpurun_out/r36_bench.err
cut -c1-330 gpurun_out/r36_bench.json
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"wilson_hop_eo" \
    --launch-skip 40 --launch-count 4 -o gpurun_out/r36_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r36_ncu_full.log 2>&1
ncu -i gpurun_out/r36_full.ncu-rep --page raw --csv > gpurun_out/r36_full_raw.csv 2>/dev/null
for k in 384 512; do
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --probes $k > gpurun_out/r36_bench_k$k.json 2> gpurun_out/r36_bench_k$k.err; cut -c1-260 gpurun_out/r36_bench_k$k.json
done
