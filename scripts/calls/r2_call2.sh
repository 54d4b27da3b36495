set -x
python -m pytest tests/test_parity_gpu.py -m gpu -x -q --durations=8 > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest2.log
tail -5 gpurun_out/r2_pytest2.log
python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err
MMK_PLACE_TB=32 python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/r2_bench2_tb32.json 2> gpurun_out/r2_bench2_tb32.err
python -m pytest tests/test_full_parity_gpu.py tests/test_full_size_gpu.py -m gpu -x -q --durations=8 > gpurun_out/r2_pytest2b.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest2b.log
tail -5 gpurun_out/r2_pytest2b.log
export MMK_PROFILE_BGZF=0
python scripts/profile_step.py 0.1 > gpurun_out/r2_ps2_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_place_block|k_place_big|k_mol_text|k_tile_mol0_big" -c 4 -o gpurun_out/r2_place_fused python scripts/profile_step.py 0.1 > gpurun_out/r2_ncu3.log 2>&1
cat gpurun_out/r2_ps2_plain.log
python -c "
import json
for f in ('r2_bench2','r2_bench2_tb32'):
    d=json.load(open('gpurun_out/'+f+'.json')); print(f, d['ms_per_step'], d['phase_ms_per_step'], d['roofline']['frac'])
"
