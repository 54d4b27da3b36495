set -x
export MMK_PROFILE_BGZF=0
for v in 1 2 4 8; do
  MMK_NVCC_FLAGS="-DMMK_EMIT_ILP=$v" python mimick_b200/build.py --force > /dev/null 2>&1
  python scripts/profile_step.py 0.25 > gpurun_out/r2_ilp$v.log 2>&1
  echo "ILP $v: $(tail -1 gpurun_out/r2_ilp$v.log | grep -o 'emit_ms=[0-9.]*')"
done
python mimick_b200/build.py --force > /dev/null 2>&1
python scripts/profile_step.py 0.25 > gpurun_out/r2_ilp_default.log 2>&1
echo "default again: $(tail -1 gpurun_out/r2_ilp_default.log | grep -o 'emit_ms=[0-9.]*')"
