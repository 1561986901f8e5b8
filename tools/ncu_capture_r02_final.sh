# final capture recipe of round 2 (one GPU): GPU suite, smoke, bench, ncu launch list of a short bench run, full-set
# captures of the compositing and sampler kernels at 65 536 rays (tools/kernel_probe.py) and of the MLP kernels of one
# train step.  Every ncu pass runs only after the same command exited 0 without ncu.
set -x
TAG=${TAG:-r02_bf}
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --sustained-s 0"
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; tail -2 gpurun_out/${TAG}_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -3 gpurun_out/${TAG}_smoke.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 400 gpurun_out/${TAG}_bench.json
$B > gpurun_out/${TAG}_plain_bench.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/${TAG}_ncu1.log 2>&1
C="python tools/kernel_probe.py --only composite"
$C > gpurun_out/${TAG}_composite_probe.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:composite -s 5 -c 1 -f -o gpurun_out/${TAG}_prof_comp_full $C > gpurun_out/${TAG}_ncu2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:composite -s 18 -c 1 -f -o gpurun_out/${TAG}_prof_comp_lean $C > gpurun_out/${TAG}_ncu3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:composite -s 31 -c 1 -f -o gpurun_out/${TAG}_prof_comp_bwd $C > gpurun_out/${TAG}_ncu4.log 2>&1
S="python tools/kernel_probe.py --only sampler"
$S > gpurun_out/${TAG}_sampler_probe.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sample_pdf_fwd -s 16 -c 1 -f -o gpurun_out/${TAG}_prof_samp_fwd $S > gpurun_out/${TAG}_ncu5.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sample_pdf_bwd -s 16 -c 1 -f -o gpurun_out/${TAG}_prof_samp_bwd $S > gpurun_out/${TAG}_ncu6.log 2>&1
$B > gpurun_out/${TAG}_plain_bench2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:mlp_tc -s 14 -c 7 -f -o gpurun_out/${TAG}_prof_mlp $B > gpurun_out/${TAG}_ncu7.log 2>&1
ls -la gpurun_out/ | tail -12
