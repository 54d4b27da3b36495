set -x
python -m pytest tests/test_parity_gpu.py -m gpu -x -q --durations=5 > gpurun_out/r2_pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest4.log
tail -5 gpurun_out/r2_pytest4.log
MMK_PLACE_MODE=fused python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/r2_bench4_f6.json 2> gpurun_out/r2_bench4_f6.err
MMK_PLACE_MODE=fused MMK_PLACE_MINB=4 python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/r2_bench4_f4.json 2> gpurun_out/r2_bench4_f4.err
MMK_PLACE_MODE=fused MMK_PLACE_WAVES=4 python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/r2_bench4_w4.json 2> gpurun_out/r2_bench4_w4.err
python -c "
import json
for f in ('r2_bench4_f6','r2_bench4_f4','r2_bench4_w4'):
    try:
        d=json.load(open('gpurun_out/'+f+'.json')); print(f, d['ms_per_step'], d['phase_ms_per_step'], d['roofline']['frac'], d['plan'])
    except Exception as e: print(f, 'failed', e)
"
MMK_PLACE_MODE=fused python bench.py --workload c4 --steps 3 --warmup 2 --e2e-steps 2 --e2e-warmup 1 --e2e-text-steps 0 --cpu-pairs 500000 > gpurun_out/r2_bench4_c4.json 2> gpurun_out/r2_bench4_c4.err
tail -3 gpurun_out/r2_bench4_c4.err; head -c 3000 gpurun_out/r2_bench4_c4.json
export MMK_PROFILE_BGZF=0
MMK_PLACE_MODE=fused python scripts/profile_step.py 0.1 > gpurun_out/r2_ps4_plain.log 2>&1 && \
MMK_PLACE_MODE=fused ncu --set full --clock-control none --import-source on -k regex:"k_place_mol|k_tile_flags" -c 2 -o gpurun_out/r2_place_mol python scripts/profile_step.py 0.1 > gpurun_out/r2_ncu6.log 2>&1
cat gpurun_out/r2_ps4_plain.log
