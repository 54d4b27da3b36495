set -x
python -m pytest tests/test_parity_gpu.py -m gpu -x -q > gpurun_out/r2_pytest27.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest27.log
tail -4 gpurun_out/r2_pytest27.log
python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/r2_bench27.json 2> gpurun_out/r2_bench27.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench27.json')); print('RES c5', d['ms_per_step'], d['phase_ms_per_step'], d['roofline']['frac'])
"
python scripts/rank_step.py 8 0 > gpurun_out/r2_rank27.log 2>&1; tail -1 gpurun_out/r2_rank27.log
MMK_PLAN_MODE=kernels python scripts/rank_step.py 8 0 > gpurun_out/r2_rank27k.log 2>&1; tail -1 gpurun_out/r2_rank27k.log
