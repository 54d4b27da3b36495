set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --e2e-steps 1 --e2e-warmup 1 --e2e-text-steps 0 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; echo "bench2 rc=$?"
tail -1 gpurun_out/r2_bench_2gpu.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('RES2', d['ms_per_step'], d['phase_ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks'])
"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_ref_2gpu.json 2> gpurun_out/r2_ref_2gpu.err; echo "ref2 rc=$?"; tail -1 gpurun_out/r2_ref_2gpu.json | cut -c1-200
