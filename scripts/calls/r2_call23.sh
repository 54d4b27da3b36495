set -x
python -m pytest tests/test_parity_gpu.py -m gpu -x -q > gpurun_out/r2_pytest23.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest23.log
tail -4 gpurun_out/r2_pytest23.log
