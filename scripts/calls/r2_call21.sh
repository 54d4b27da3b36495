set -x
python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "placement or ci_single or formats or long_contig" > gpurun_out/r2_pytest21.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest21.log
tail -4 gpurun_out/r2_pytest21.log
python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/r2_bench21.json 2> gpurun_out/r2_bench21.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench21.json')); print('RES c5', d['ms_per_step'], d['phase_ms_per_step'], d['roofline']['frac'])
"
python bench.py --workload c4 --steps 3 --warmup 3 --e2e-steps 2 --e2e-warmup 1 --e2e-text-steps 0 --cpu-pairs 500000 > gpurun_out/r2_bench21_c4.json 2> gpurun_out/r2_bench21_c4.err; echo "c4 rc=$?"
tail -3 gpurun_out/r2_bench21_c4.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench21_c4.json')); print('RES c4', d['value'], d['ms_per_step'], d.get('phase_ms_per_step'), d['roofline']['frac'], d['e2e']['value'], d['config'].get('haplotype_construction'))
"
