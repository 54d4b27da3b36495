set -x
./scripts/micro/tile_store > gpurun_out/r2_tile_store.log 2>&1; cat gpurun_out/r2_tile_store.log
python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "error" --durations=3 > gpurun_out/r2_pytest10.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest10.log
tail -30 gpurun_out/r2_pytest10.log
