set -x
python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_final.json')); print('RES c5', d['ms_per_step'], d['phase_ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['e2e_text']['value'] if d.get('e2e_text') else None, d['cpu_baseline']['value'], d['clocks'])
"
timeout 600 python -m pytest tests -m gpu -x -q --durations=10 > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_final.log
tail -16 gpurun_out/r2_pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -2 gpurun_out/r2_smoke.log
timeout 200 python scripts/file_e2e.py 0.1 > gpurun_out/r2_file_e2e.json 2> gpurun_out/r2_file_e2e.err; echo "file rc=$?"; tail -1 gpurun_out/r2_file_e2e.json | cut -c1-700
