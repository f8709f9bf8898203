# round 2, run 28 (1 GPU): operator memory from the stream-ordered pool: GPU tier, set-up profiles 128^2 / 512^2 / 1024^2, bench twice
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2_28_pytest.log
grep -E "Error|passed|failed" gpurun_out/r2_28_pytest.log | head -8 | cut -c1-400
timeout 300 python profiles/profile_setup.py --lines 30 > gpurun_out/r2_28_profile_setup_128.log 2>&1
grep "setup wall" gpurun_out/r2_28_profile_setup_128.log
timeout 600 python profiles/profile_setup.py --L 512 --lines 40 > gpurun_out/r2_28_profile_setup_512.log 2>&1
grep -E "setup wall|Error|error" gpurun_out/r2_28_profile_setup_512.log | cut -c1-300
timeout 900 python profiles/profile_setup.py --L 1024 --lines 40 > gpurun_out/r2_28_profile_setup_1024.log 2>&1
grep -E "setup wall|Error|error" gpurun_out/r2_28_profile_setup_1024.log | cut -c1-300
for i in 1 2; do
timeout 600 python bench.py --no-cpu-baseline --no-experiment > gpurun_out/r2_28_bench_$i.json 2> gpurun_out/r2_28_bench_$i.err
python -c "
import json,sys
d=json.loads(open('gpurun_out/r2_28_bench_$i.json').read()); print(d['value'], d['e2e']['value'], d['fgmres_iters'], d['gpu_launches'], d.get('setup_s'))"
done
