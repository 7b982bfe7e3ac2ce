set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_r01d.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/launches_r01d.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:route_hist_level --launch-skip 25 -c 1 -f -o gpurun_out/route_r01g python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/route_r01g.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:score_level --launch-skip 28 -c 1 -f -o gpurun_out/score_r01g python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/score_r01g.log 2>&1
ls -la gpurun_out
