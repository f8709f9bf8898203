set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r21_pytest.log
for d in 24 32 48; do
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --degree $d > gpurun_out/r21_bench_d$d.json 2> gpurun_out/r21_bench_d$d.err
done
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --precond reference > gpurun_out/r21_bench_ref.json 2> gpurun_out/r21_bench_ref.err
tail -3 gpurun_out/r21_pytest.log; cat gpurun_out/r21_bench_d*.json | cut -c1-400
