# run 39: final state of the round: GPU tests, smoke, bench (tcgen05 grid rasterised so that the column tiles of a matrix slab run together)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r39_pytest.log
tail -4 gpurun_out/r39_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r39_smoke.log 2>&1; tail -1 gpurun_out/r39_smoke.log
python bench.py --steps 8 --warmup 3 > gpurun_out/r39_bench.json 2> gpurun_out/r39_bench.err || tail -5 gpurun_out/r39_bench.err
cut -c1-330 gpurun_out/r39_bench.json
python bench.py --steps 4 --warmup 3 --no-cpu-baseline --probes 256 > gpurun_out/r39_bench_k256.json 2> gpurun_out/r39_bench_k256.err; cut -c1-260 gpurun_out/r39_bench_k256.json
