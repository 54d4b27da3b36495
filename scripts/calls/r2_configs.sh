set -x
for wl in c4 c3 c2; do
  timeout 200 python bench.py --workload $wl --steps 5 --warmup 3 --e2e-steps 2 --e2e-warmup 1 --e2e-text-steps 0 --no-cpu > gpurun_out/r2_final_$wl.json 2> gpurun_out/r2_final_$wl.err; echo "$wl rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/r2_final_$wl.json')); print('RES $wl', d['value'], d['ms_per_step'], d['phase_ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['fastq_gb_per_s'], d.get('haplotype_construction'))
"
done
