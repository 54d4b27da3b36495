set -x
for v in "MMK_EMIT_MODE=warp MMK_EMIT_ILP=4" "MMK_EMIT_MODE=cta" "MMK_EMIT_MODE=warp MMK_EMIT_ILP=3"; do
  env $v timeout 300 python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/r2_bench17_tmp.json 2> gpurun_out/r2_bench17_tmp.err
  python -c "
import json
d=json.load(open('gpurun_out/r2_bench17_tmp.json')); print('RES $v', d['ms_per_step'], d['phase_ms_per_step'], d['roofline']['frac'])
"
done
MMK_EMIT_MODE=warp MMK_EMIT_ILP=4 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "ci_single or circular or formats" > gpurun_out/r2_pytest17.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest17.log
tail -4 gpurun_out/r2_pytest17.log
