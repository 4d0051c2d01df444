#!/usr/bin/env bash
# round-2 final ncu evidence: launch list (time + DRAM bytes) of one eager step, and --set full of the new GEMM families
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --csv --log-file gpurun_out/launches_r02_final.csv python tools/ncu_step.py > gpurun_out/ncu_step.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_r02_final.csv gpurun_out/launches_r02_final.txt "round 2 final (LayerNorm fold, epilogue families)" | tail -25
ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/ln_gemm_r02 -f python tools/ncu_ln_gemm.py > gpurun_out/ncu_ln.log 2>&1
ncu -i gpurun_out/ln_gemm_r02.ncu-rep --page raw --csv > gpurun_out/ln_gemm_r02_raw.csv 2>/dev/null
ls -la gpurun_out/*.ncu-rep gpurun_out/ln_gemm_r02_raw.csv
