set -x
L=mimick_b200/libmimick_b200.so
for v in 0 1 0 1; do
  cp scripts/micro/variants/lib_hdr$v.so $L
  python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/r2_hdr$v.json 2> gpurun_out/r2_hdr$v.err
  python -c "
import json
d=json.load(open('gpurun_out/r2_hdr$v.json')); print('HDR $v', d['ms_per_step'], d['phase_ms_per_step'], d['roofline']['frac'])
"
done
cp scripts/micro/variants/lib_hdr1.so $L
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_hdr1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_hdr1.log
tail -4 gpurun_out/r2_pytest_hdr1.log
