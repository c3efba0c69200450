set -x
python bench.py > gpurun_out/bench_s3_b.json 2> gpurun_out/bench_s3_b.err; tail -c 2500 gpurun_out/bench_s3_b.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r01_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1
for k in metropolis_reg adjoint_warp pgrad_binned_kernel eloc2; do
  ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/r01_$k python scripts/prof_step.py --walkers 9472 --iters 1 > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out
