# round-2 profiling pass (GPU box): plain runs first, then the launch list and two small full captures -- numbers printed under ncu are never
# bench values.  The .ncu-rep files (about 1 MB per launch) are exported to CSV on the box and removed: gpurun copies back at most 64 MiB.
set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-legs > gpurun_out/r2_prof_plain.json 2> gpurun_out/r2_prof_plain.err || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-legs > gpurun_out/r2_ncu_launch.log 2>&1
python scripts/ncu_target.py > gpurun_out/r2_target_plain.log 2>&1 || exit 1
timeout 400 ncu --set full --clock-control none --profile-from-start off -k regex:'k_|sor_' -c 22 -o gpurun_out/r2_full_a -f python scripts/ncu_target.py > gpurun_out/r2_ncu_full_a.log 2>&1
timeout 400 ncu --set full --clock-control none --profile-from-start off -k regex:'sor_rb_tiled|sor_gs_tiled_kernel<\(bool\)0|k_advect|k_extrapolate' -c 7 -o gpurun_out/r2_full_b -f python scripts/ncu_target.py > gpurun_out/r2_ncu_full_b.log 2>&1
for x in a b; do ncu -i gpurun_out/r2_full_$x.ncu-rep --page raw --csv > gpurun_out/r2_full_${x}_raw.csv 2>/dev/null; rm -f gpurun_out/r2_full_$x.ncu-rep; done
tail -2 gpurun_out/r2_ncu_full_a.log gpurun_out/r2_ncu_full_b.log; ls -la gpurun_out/; du -sh gpurun_out
