set -x
python -m pytest tests/test_parity_gpu.py -m gpu -x -q --durations=5 > gpurun_out/r2_pytest7.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest7.log
tail -25 gpurun_out/r2_pytest7.log
python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/r2_bench7_new.json 2> gpurun_out/r2_bench7_new.err
MMK_EMIT_MODE=cta python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/r2_bench7_cta.json 2> gpurun_out/r2_bench7_cta.err
python -c "
import json
for f in ('r2_bench7_new','r2_bench7_cta'):
    try:
        d=json.load(open('gpurun_out/'+f+'.json')); print(f, d['ms_per_step'], d['phase_ms_per_step'], d['roofline']['frac'], d.get('plan'))
    except Exception as e: print(f, 'failed', e)
"
tail -3 gpurun_out/r2_bench7_new.err
export MMK_PROFILE_BGZF=0
python scripts/profile_step.py 0.1 > gpurun_out/r2_ps7_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_place_block|k_emit_w" -c 3 -o gpurun_out/r2_new7 python scripts/profile_step.py 0.1 > gpurun_out/r2_ncu7.log 2>&1
cat gpurun_out/r2_ps7_plain.log
