set -x
nproc; nvidia-smi -L
python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest6.log
tail -15 gpurun_out/r2_pytest6.log
python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/r2_bench6_legacy.json 2> gpurun_out/r2_bench6_legacy.err
MMK_PLACE_MODE=fused python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/r2_bench6_fused.json 2> gpurun_out/r2_bench6_fused.err
python -c "
import json
for f in ('r2_bench6_legacy','r2_bench6_fused'):
    try:
        d=json.load(open('gpurun_out/'+f+'.json')); print(f, d['ms_per_step'], d['phase_ms_per_step'], d['roofline']['frac'], d.get('plan'))
    except Exception as e: print(f, 'failed', e)
"
export MMK_PROFILE_BGZF=0
python scripts/profile_step.py 0.1 > gpurun_out/r2_ps6_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches6_legacy.csv python scripts/profile_step.py 0.1 > gpurun_out/r2_ncu6a.log 2>&1
MMK_PLACE_MODE=fused python scripts/profile_step.py 0.1 > gpurun_out/r2_ps6_fused.log 2>&1 && \
MMK_PLACE_MODE=fused ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches6_fused.csv python scripts/profile_step.py 0.1 > gpurun_out/r2_ncu6b.log 2>&1
cat gpurun_out/r2_ps6_plain.log gpurun_out/r2_ps6_fused.log
