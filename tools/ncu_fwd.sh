set -x
P="python tools/kernel_probe.py --rays 2048 --only mlp"
$P > gpurun_out/plain_probe_fwd.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mlp_tc_fwd -s 4 -c 1 -o gpurun_out/prof_fwd2 $P > gpurun_out/ncu_fwd2.log 2>&1
ls -la gpurun_out/*.ncu-rep
