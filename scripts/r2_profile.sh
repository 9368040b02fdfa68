# round-2 profiling pass (GPU box): plain run first, then the launch list and one full capture -- numbers printed under ncu are never bench values
set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-legs > gpurun_out/r2_prof_plain.json 2> gpurun_out/r2_prof_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-legs > gpurun_out/r2_ncu_launch.log 2>&1
python scripts/ncu_target.py > gpurun_out/r2_target_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'k_|sor_' -c 400 -o gpurun_out/r2_full -f python scripts/ncu_target.py > gpurun_out/r2_ncu_full.log 2>&1
tail -2 gpurun_out/r2_ncu_full.log; ls -la gpurun_out/r2_full.ncu-rep
