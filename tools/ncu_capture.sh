# round-1 capture recipe: launch list of one bench run + full-set captures of the MLP kernels
set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$B > gpurun_out/plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu1.log 2>&1
$B > gpurun_out/plain_bench2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mlp_tc -s 12 -c 6 -o gpurun_out/prof_mlp $B > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out/
