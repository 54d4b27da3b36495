set -x
# BGZF member size: one tile (32 records) per member against two (MMK_BGZF_TILES=2), same shard
python scripts/profile_step.py 0.1 > gpurun_out/r2_bgzf_g1.log 2>&1; echo "g1 rc=$?"; tail -1 gpurun_out/r2_bgzf_g1.log
MMK_BGZF_TILES=2 python scripts/profile_step.py 0.1 > gpurun_out/r2_bgzf_g2.log 2>&1; echo "g2 rc=$?"; tail -1 gpurun_out/r2_bgzf_g2.log
export MMK_PROFILE_BGZF=0
python scripts/profile_step.py 0.1 > gpurun_out/r2_prof_plain.log 2>&1; echo "plain rc=$?"; tail -1 gpurun_out/r2_prof_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_final_launches.csv python scripts/profile_step.py 0.1 > gpurun_out/r2_ncu_list.log 2>&1; echo "list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"k_place_block\$|k_emit_w" -c 2 -o gpurun_out/r2_final_full python scripts/profile_step.py 0.1 > gpurun_out/r2_ncu_full.log 2>&1; echo "full rc=$?"
ls -la gpurun_out/*.ncu-rep
