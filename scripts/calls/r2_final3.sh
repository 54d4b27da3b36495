set -x
timeout 700 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/r2_pytest_final3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_final3.log
tail -20 gpurun_out/r2_pytest_final3.log
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "ref rc=$?"; cut -c1-600 gpurun_out/r2_bench_reference.json
