set -x
for v in "MMK_EMIT_MODE=cta" "MMK_EMIT_ILP=1" "MMK_EMIT_ILP=3"; do
  env $v python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/r2_bench12_tmp.json 2> gpurun_out/r2_bench12_tmp.err
  python -c "
import json
d=json.load(open('gpurun_out/r2_bench12_tmp.json')); print('$v', d['ms_per_step'], d['phase_ms_per_step'], d['roofline']['frac'])
"
done
