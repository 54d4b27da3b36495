set -x
python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "long_contig or placement or ci_single" > gpurun_out/r2_pytest20.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest20.log
tail -4 gpurun_out/r2_pytest20.log
python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/r2_bench20.json 2> gpurun_out/r2_bench20.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench20.json')); print('RES', d['ms_per_step'], d['phase_ms_per_step'], d['roofline']['frac'])
"
export MMK_PROFILE_BGZF=0
python scripts/profile_step.py 1.0 > gpurun_out/r2_ps20_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_plan|k_place|k_mol|k_tile|k_scan|k_group|k_fill|k_shard|k_chunk|k_local" -c 1200 --csv --log-file gpurun_out/r2_launches20_full.csv python scripts/profile_step.py 1.0 > gpurun_out/r2_ncu20.log 2>&1
cat gpurun_out/r2_ps20_plain.log
