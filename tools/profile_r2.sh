#!/bin/sh
# ncu evidence for profiles/ (round 2): launch list of whole steps + a `--set full` capture of the
# dominant kernel + lighter section captures of one launch of every other rated kernel.  Run under
# gpurun (one GPU).  Only CSV pages and the top kernel's report come back (gpurun returns <= 64 MiB).
CMD="python bench.py --steps 2 --warmup 3 --spinup 3 --no-cpu"
O=gpurun_out
SEC="--section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section LaunchStats"
$CMD > $O/profile_plain.log 2>&1 || { echo "plain run failed"; tail -5 $O/profile_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 4700 -c 1600 --csv --log-file $O/ncu_r2_launches.csv $CMD > $O/ncu_r2_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'^vk_ell2_row_op' -s 60 -c 3 -o $O/ncu_r2_ell2 -f $CMD > $O/ncu_r2_ell2.log 2>&1
ncu -i $O/ncu_r2_ell2.ncu-rep --page raw --csv > $O/ncu_r2_ell2_raw.csv
ncu $SEC --clock-control none -k regex:'^(k_grad_scalar|k_alpha_flux|k_mules_setup|k_mules_cell|k_mules_face)' -s 70 -c 5 -o /tmp/a -f $CMD > $O/ncu_r2_alpha.log 2>&1
ncu -i /tmp/a.ncu-rep --page raw --csv > $O/ncu_r2_alpha_raw.csv
ncu $SEC --clock-control none -k regex:'^(k_U_recon|k_HbyA|k_phiHbyA|k_mom_face|k_grad_U|k_flux|k_Uf)' -s 28 -c 9 -o /tmp/m -f $CMD > $O/ncu_r2_mom.log 2>&1
ncu -i /tmp/m.ncu-rep --page raw --csv > $O/ncu_r2_mom_raw.csv
ncu $SEC --clock-control none -k regex:'^(vk_tail|vk_ellc_row_op|k_spmv_dot_ell2|vk_restrict|k_update_xr)' -s 40 -c 8 -o /tmp/s -f $CMD > $O/ncu_r2_solver.log 2>&1
ncu -i /tmp/s.ncu-rep --page raw --csv > $O/ncu_r2_solver_raw.csv
ls -la $O | tail -12; du -sh $O
