# Round-2 measurement batch (one B200): tests, bench line, launch list, full captures of the E_loc sweep, its finale and the
# second-tier kernels, the other BASELINE configurations.  Outputs under gpurun_out/ (copied to profiles/ by hand).
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02_pytest_final.txt; cat gpurun_out/r02_pytest_final.txt
timeout 400 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; cut -c1-600 gpurun_out/r02_bench_final.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1
for k in eloc5_kernel eloc_finale_warp flow_warp_kernel pgrad_binned_kernel adjoint_warp_kernel metropolis_reg_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/r02_$k python scripts/prof_step.py --walkers 9472 --iters 1 > gpurun_out/ncu_r02_$k.log 2>&1
done
rm -f gpurun_out/r02_config_lines.jsonl
for c in readme_finiteT strong_coupling slater_sweep; do timeout 300 python bench.py --config $c 2>/dev/null >> gpurun_out/r02_config_lines.jsonl; done
wc -c gpurun_out/r02_config_lines.jsonl
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2>/dev/null; cut -c1-400 gpurun_out/r02_bench_reference.json
python scripts/dev_step_parts.py > gpurun_out/r02_step_parts.txt; cat gpurun_out/r02_step_parts.txt
python scripts/dev_e5_timing.py > gpurun_out/r02_e5_cycles.txt 2>&1 || true
FF_DEV_LIB=libff_onecta.so python scripts/dev_e5_timing.py 1184 > gpurun_out/r02_e5_cycles_onecta.txt 2>&1 || true
