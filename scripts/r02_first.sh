set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_pytest_a.txt; cat gpurun_out/r02_pytest_a.txt
python bench.py > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; tail -c 3000 gpurun_out/r02_bench_a.json; tail -5 gpurun_out/r02_bench_a.err
./scripts/ubench/dmma_shapes > gpurun_out/r02_dmma_shapes.txt 2>&1; cat gpurun_out/r02_dmma_shapes.txt
for k in 0 5 8; do
  for w in 148 2368; do
    FF_TIMING_LIB=libff_timing_$k.so python scripts/dev_phase_timing2.py $w >> gpurun_out/r02_phase_eloc2.txt 2>&1
  done
done
cat gpurun_out/r02_phase_eloc2.txt
