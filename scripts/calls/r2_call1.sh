set -x
python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest1.log
python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
export MMK_PROFILE_BGZF=0
python scripts/profile_step.py 0.1 > gpurun_out/r2_ps_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_insert_draws|k_place_chain|k_tile_prepare|k_mol_text" -c 4 -o gpurun_out/r2_place_old python scripts/profile_step.py 0.1 > gpurun_out/r2_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_plan_molecules|k_plan_cut|k_plan_sums" -s 30 -c 3 -o gpurun_out/r2_plan_old python scripts/profile_step.py 0.1 > gpurun_out/r2_ncu2.log 2>&1
tail -3 gpurun_out/r2_pytest1.log; cat gpurun_out/r2_bench1.json | head -c 1500
