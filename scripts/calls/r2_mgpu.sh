set -x
N=$1; WL=${2:-c5}
nvidia-smi -L | head -8; nproc
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $WL $3 > gpurun_out/r2_bench_${WL}_${N}gpu.json 2> gpurun_out/r2_bench_${WL}_${N}gpu.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_${WL}_${N}gpu.err
python -c "
import json
d=[json.loads(l) for l in open('gpurun_out/r2_bench_${WL}_${N}gpu.json') if l.startswith('{')][-1]; print('RES', '$WL', $N, d['value'], d['ms_per_step'], d.get('phase_ms_per_step'), d['roofline']['frac'], d['e2e']['value'], d['e2e'].get('pcie_gb_per_s'), d['e2e'].get('host_phase_ms_per_step'))
"
