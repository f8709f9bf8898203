# run 23: batch-size sweep of the geometric-preconditioner configuration
mkdir -p gpurun_out
for k in 128 384 512 768; do
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --probes $k > gpurun_out/r23_bench_k$k.json 2> gpurun_out/r23_bench_k$k.err
  cut -c1-260 gpurun_out/r23_bench_k$k.json
done
