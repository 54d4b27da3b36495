set -x
python -m pytest tests/test_parity_gpu.py -m gpu -x -q --durations=8 > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest3.log
tail -5 gpurun_out/r2_pytest3.log
python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
MMK_PLACE_MODE=fused python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/r2_bench3_f64.json 2> gpurun_out/r2_bench3_f64.err
MMK_PLACE_MODE=fused MMK_PLACE_TB=32 python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/r2_bench3_f32.json 2> gpurun_out/r2_bench3_f32.err
python -c "
import json
for f in ('r2_bench3','r2_bench3_f64','r2_bench3_f32'):
    try:
        d=json.load(open('gpurun_out/'+f+'.json')); print(f, d['ms_per_step'], d['phase_ms_per_step'], d['roofline']['frac'], d['plan'])
    except Exception as e: print(f, 'failed', e)
"
python -m pytest tests/test_full_parity_gpu.py tests/test_full_size_gpu.py -m gpu -x -q --durations=8 > gpurun_out/r2_pytest3b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest3b.log
tail -5 gpurun_out/r2_pytest3b.log
export MMK_PROFILE_BGZF=0
MMK_PLACE_MODE=fused python scripts/profile_step.py 0.1 > gpurun_out/r2_ps3_plain.log 2>&1 && \
MMK_PLACE_MODE=fused ncu --set full --clock-control none --import-source on -k regex:"k_place_block|k_place_big" -c 2 -o gpurun_out/r2_place_fused2 python scripts/profile_step.py 0.1 > gpurun_out/r2_ncu4.log 2>&1
cat gpurun_out/r2_ps3_plain.log
python scripts/profile_step.py 0.1 > gpurun_out/r2_ps3b_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_insert_draws|k_place_chain|k_tile_prepare|k_emit" -c 4 -o gpurun_out/r2_place_legacy2 python scripts/profile_step.py 0.1 > gpurun_out/r2_ncu5.log 2>&1
cat gpurun_out/r2_ps3b_plain.log
