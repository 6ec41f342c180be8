set -x
CMD="python bench.py --steps 10 --warmup 3 --no-cpu --e2e-steps 2 --no-secondary --sustained-s 0"
$CMD > gpurun_out/r2n_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c_launches.csv $CMD > gpurun_out/r2n_ncu1.log 2>&1
CMD2="python bench.py --steps 5 --warmup 3 --no-cpu --e2e-steps 2 --no-secondary --sustained-s 0"
$CMD2 > gpurun_out/r2n_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 4 -c 1 -o gpurun_out/r02c_step_kernel $CMD2 > gpurun_out/r2n_ncu2.log 2>&1
ls -la gpurun_out/r02c*
