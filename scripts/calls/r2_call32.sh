set -x
python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/r2_bench32.json 2> gpurun_out/r2_bench32.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench32.json')); print('RES c5', d['ms_per_step'], d['phase_ms_per_step'], d['roofline']['frac'])
"
export MMK_PROFILE_BGZF=0
for v in 592 2368; do
  MMK_NVCC_FLAGS="-DMMK_EMIT_AHEAD=$v" python mimick_b200/build.py --force > /dev/null 2>&1
  python scripts/profile_step.py 0.25 > gpurun_out/r2_ahead$v.log 2>&1
  echo "AHEAD $v: $(tail -1 gpurun_out/r2_ahead$v.log | grep -o 'emit_ms=[0-9.]*')"
done
