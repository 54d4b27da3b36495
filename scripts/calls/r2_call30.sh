set -x
python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/r2_bench30.json 2> gpurun_out/r2_bench30.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench30.json')); print('RES c5', d['ms_per_step'], d['phase_ms_per_step'], d['roofline']['frac'])
"
python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "sharding or ci_single or bgzf_members" > gpurun_out/r2_pytest30.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest30.log
tail -3 gpurun_out/r2_pytest30.log
export MMK_PROFILE_BGZF=0
python scripts/profile_step.py 0.1 > gpurun_out/r2_ps30_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_place_block\$|k_emit_w" -c 2 -o gpurun_out/r2_final30 python scripts/profile_step.py 0.1 > gpurun_out/r2_ncu30.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_launches30.csv python scripts/profile_step.py 0.1 > gpurun_out/r2_ncu30b.log 2>&1
cat gpurun_out/r2_ps30_plain.log
