set -x
timeout 700 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2_pytest_final5.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_final5.log
tail -14 gpurun_out/r2_pytest_final5.log
timeout 300 python -m pytest tests/test_full_parity_gpu.py -m gpu -x -q -k "c2 or c5" > gpurun_out/r2_pytest_alone.log 2>&1; echo "alone rc=$?"; tail -2 gpurun_out/r2_pytest_alone.log
