# round-1 capture recipe: launch list of one bench run + full-set captures of the MLP kernels of one train step.
# Every ncu pass runs only after the same command exited 0 without ncu.
set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
TAG=${TAG:-r01_i}
$B > gpurun_out/${TAG}_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/${TAG}_ncu1.log 2>&1
$B > gpurun_out/${TAG}_plain_bench2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mlp_tc -s 14 -c 7 -f -o gpurun_out/${TAG}_prof_mlp $B > gpurun_out/${TAG}_ncu2.log 2>&1
ls -la gpurun_out/ | tail -5
