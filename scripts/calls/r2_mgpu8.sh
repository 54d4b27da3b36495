set -x
nvidia-smi -L | wc -l; nproc
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --workload c5 --steps 5 --warmup 3 --e2e-steps 2 --e2e-warmup 1 --e2e-text-steps 0 > gpurun_out/r2_bench_c5_8gpu_final.json 2> gpurun_out/r2_bench_c5_8gpu_final.err; echo "bench rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 scripts/d2h_probe.py gpurun_out/r2_d2h_probe_8gpu.json > gpurun_out/r2_d2h_probe_8gpu.log 2>&1; echo "probe rc=$?"
tail -1 gpurun_out/r2_d2h_probe_8gpu.log | cut -c1-600
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/r2_bench_c5_8gpu_final.json') if l.startswith('{')][-1]; print('RES c5 8', d['value'], d['ms_per_step'], d.get('phase_ms_per_step'), d['roofline']['frac'], d['e2e']['value'], d['e2e'].get('pcie_gb_per_s'))
"
