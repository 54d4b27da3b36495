set -x
python -m pytest tests/test_parity_gpu.py -m gpu -x -q > gpurun_out/r2_pytest13.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest13.log
tail -6 gpurun_out/r2_pytest13.log
for v in "MMK_EMIT_MODE=persistent" "MMK_EMIT_MODE=cta"; do
  env $v python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/r2_bench13_tmp.json 2> gpurun_out/r2_bench13_tmp.err
  python -c "
import json
d=json.load(open('gpurun_out/r2_bench13_tmp.json')); print('RES $v', d['ms_per_step'], d['phase_ms_per_step'], d['roofline']['frac'])
"
done
export MMK_PROFILE_BGZF=0
python scripts/profile_step.py 0.1 > gpurun_out/r2_ps13_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_emit_p" -c 1 -o gpurun_out/r2_new13 python scripts/profile_step.py 0.1 > gpurun_out/r2_ncu13.log 2>&1
cat gpurun_out/r2_ps13_plain.log
