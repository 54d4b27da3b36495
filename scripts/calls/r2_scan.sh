set -x
L=mimick_b200/libmimick_b200.so
for v in 0 1 0 1; do
  cp scripts/micro/variants/lib_scan$v.so $L
  python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/r2_scan$v.json 2> gpurun_out/r2_scan$v.err
  python -c "
import json
d=json.load(open('gpurun_out/r2_scan$v.json')); print('SCAN $v', d['ms_per_step'], d['phase_ms_per_step'], d['roofline']['frac'])
"
done
cp scripts/micro/variants/lib_scan1.so $L
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_scan1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_scan1.log
tail -4 gpurun_out/r2_pytest_scan1.log
