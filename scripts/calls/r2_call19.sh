set -x
python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/r2_pytest19.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest19.log
tail -12 gpurun_out/r2_pytest19.log
python bench.py > gpurun_out/r2_bench19_default.json 2> gpurun_out/r2_bench19_default.err; echo "bench rc=$?"
tail -2 gpurun_out/r2_bench19_default.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench19_default.json')); print('RES', d['ms_per_step'], d['phase_ms_per_step'], d['roofline'], d['e2e'], d.get('cpu_baseline'))
"
