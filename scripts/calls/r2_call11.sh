set -x
python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "ci_single or circular or c1_multi" > gpurun_out/r2_pytest11.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest11.log
tail -4 gpurun_out/r2_pytest11.log
python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/r2_bench11_la.json 2> gpurun_out/r2_bench11_la.err
python -c "
import json
for f in ('r2_bench11_la',):
    try:
        d=json.load(open('gpurun_out/'+f+'.json')); print(f, d['ms_per_step'], d['phase_ms_per_step'], d['roofline']['frac'], d.get('plan'))
    except Exception as e: print(f, 'failed', e)
"
tail -3 gpurun_out/r2_bench11_la.err
export MMK_PROFILE_BGZF=0
python scripts/profile_step.py 0.1 > gpurun_out/r2_ps11_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_emit_w" -c 1 -o gpurun_out/r2_new11 python scripts/profile_step.py 0.1 > gpurun_out/r2_ncu11.log 2>&1
cat gpurun_out/r2_ps11_plain.log
