set -u
# round-2 final-state profile pass (run on the GPU box through gpurun); .ncu-rep files stay in gpurun_out/
python tools/prof_step.py > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2b_launches_bench.csv python bench.py --steps 64 --warmup 3 --no-cpu > gpurun_out/r2b_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"project|splat" -s 16 -c 4 -f -o gpurun_out/r2b_prof_step python tools/prof_step.py > gpurun_out/r2b_ncu_step.log 2>&1
PFUSED=1 ncu --set full --clock-control none --import-source on -k regex:"project|splat" -s 36 -c 3 -f -o gpurun_out/r2b_prof_fused python tools/prof_step.py > gpurun_out/r2b_ncu_fused.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
