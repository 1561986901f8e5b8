# round-2 capture recipe: launch list of one bench run, full-set captures of the MLP kernels of one train step, and a
# full-set capture of the layer-pipelined backward prototype (seven layer groups in one launch).
# Every ncu pass runs only after the same command exited 0 without ncu.
set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --sustained-s 0"
P="python tools/bwd_pipe_probe.py --rays 4096 --samples 128 --layers 701 --iters 1"
TAG=${TAG:-r02_n}
$B > gpurun_out/${TAG}_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/${TAG}_ncu1.log 2>&1
$B > gpurun_out/${TAG}_plain_bench2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mlp_tc -s 14 -c 7 -f -o gpurun_out/${TAG}_prof_mlp $B > gpurun_out/${TAG}_ncu2.log 2>&1
$P > gpurun_out/${TAG}_plain_pipe.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mlp_tc_bwd_pipe_kernel -s 1 -c 1 -f -o gpurun_out/${TAG}_prof_pipe $P > gpurun_out/${TAG}_ncu3.log 2>&1
ls -la gpurun_out/ | tail -8
