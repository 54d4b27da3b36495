set -x
export MMK_PROFILE_BGZF=0
for v in 0 1; do
  MMK_NVCC_FLAGS="-DMMK_EMIT_EARLY=$v" python mimick_b200/build.py --force > /dev/null 2>&1
  python scripts/profile_step.py 0.25 > gpurun_out/r2_early$v.log 2>&1
  echo "EARLY $v: $(tail -1 gpurun_out/r2_early$v.log | grep -o 'emit_ms=[0-9.]*')"
done
python mimick_b200/build.py --force > /dev/null 2>&1
python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "sharding or ci_single or bgzf_members or parameter" > gpurun_out/r2_pytest33.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest33.log
tail -3 gpurun_out/r2_pytest33.log
