#!/bin/bash
# BASELINE configs 1 and 3 as whole runs on one B200 (wall-clock of the complete driver, incl. Python start-up)
mkdir -p gpurun_out
s=$(date +%s.%N)
python -m fermiflow_b200.BetaFermionHO2D --beta 10.0 --nup 3 --Z 2.0 --deltaE 2.0 --boltzmann --iternum 1000 > gpurun_out/config1_readme_run.log 2>&1
e=$(date +%s.%N)
echo "config 1 (README finite-T run, 1000 iterations, batch 8000): $(python -c "print(round($e - $s, 1))") s wall" | tee gpurun_out/config_runs.txt
head -3 gpurun_out/config1_readme_run.log | cut -c1-200 | tee -a gpurun_out/config_runs.txt
tail -1 gpurun_out/config1_readme_run.log | cut -c1-260 | tee -a gpurun_out/config_runs.txt
s=$(date +%s.%N)
python -m fermiflow_b200.BetaFermionHO2D --beta 2.0 --nup 12 --Z 8.0 --deltaE 2.0 --boltzmann --batch 8000 --iternum 300 --nsteps 32 > gpurun_out/config3_strong_coupling.log 2>&1
e=$(date +%s.%N)
echo "config 3 (Z = 8, N = 12, 32 RK4 steps, finite T, 300 iterations, batch 8000): $(python -c "print(round($e - $s, 1))") s wall" | tee -a gpurun_out/config_runs.txt
head -3 gpurun_out/config3_strong_coupling.log | cut -c1-200 | tee -a gpurun_out/config_runs.txt
tail -1 gpurun_out/config3_strong_coupling.log | cut -c1-260 | tee -a gpurun_out/config_runs.txt
