set -x
python -m pytest tests/test_parity_gpu.py -m gpu -x -q --durations=5 > gpurun_out/r2_pytest5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest5.log
tail -5 gpurun_out/r2_pytest5.log
MMK_PLACE_MODE=fused python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/r2_bench5_a.json 2> gpurun_out/r2_bench5_a.err
MMK_PLACE_MODE=fused MMK_PLACE_CACHE=1 python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/r2_bench5_b.json 2> gpurun_out/r2_bench5_b.err
MMK_PLACE_MODE=fused MMK_PLACE_CACHE=1 MMK_EMIT_QFILL=0 python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/r2_bench5_c.json 2> gpurun_out/r2_bench5_c.err
python -c "
import json
for f in ('r2_bench5_a','r2_bench5_b','r2_bench5_c'):
    try:
        d=json.load(open('gpurun_out/'+f+'.json')); print(f, d['ms_per_step'], d['phase_ms_per_step'], d['roofline']['frac'])
    except Exception as e: print(f, 'failed', e)
"
export MMK_PROFILE_BGZF=0
MMK_PLACE_MODE=fused MMK_PLACE_CACHE=1 python scripts/profile_step.py 0.1 > gpurun_out/r2_ps5_plain.log 2>&1 && \
MMK_PLACE_MODE=fused MMK_PLACE_CACHE=1 ncu --set full --clock-control none --import-source on -k regex:"k_mol_sum|k_mol_chain|k_emit" -c 3 -o gpurun_out/r2_place_mol2 python scripts/profile_step.py 0.1 > gpurun_out/r2_ncu7.log 2>&1
cat gpurun_out/r2_ps5_plain.log
