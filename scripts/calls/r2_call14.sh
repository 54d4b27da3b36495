set -x
python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "ci_single or circular or c1_multi or formats or placement" > gpurun_out/r2_pytest14.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest14.log
tail -6 gpurun_out/r2_pytest14.log
for v in "MMK_EMIT_MODE=persistent" "MMK_EMIT_CTAS=3"; do
  env $v python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/r2_bench14_tmp.json 2> gpurun_out/r2_bench14_tmp.err
  python -c "
import json
d=json.load(open('gpurun_out/r2_bench14_tmp.json')); print('RES $v', d['ms_per_step'], d['phase_ms_per_step'], d['roofline']['frac'])
"
done
export MMK_PROFILE_BGZF=0
python scripts/profile_step.py 0.1 > gpurun_out/r2_ps14_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_emit_p" -c 1 -o gpurun_out/r2_new14 python scripts/profile_step.py 0.1 > gpurun_out/r2_ncu14.log 2>&1
cat gpurun_out/r2_ps14_plain.log
