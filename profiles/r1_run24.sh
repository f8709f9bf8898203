# run 24: fused V-cycle I/O + split-BF16 dense apply: GPU tests, bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r24_pytest.log
tail -4 gpurun_out/r24_pytest.log
python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r24_bench.json 2> gpurun_out/r24_bench.err
cut -c1-330 gpurun_out/r24_bench.json
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r24_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r24_ncu_bench.log 2>&1
