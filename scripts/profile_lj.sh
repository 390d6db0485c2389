#!/bin/bash
# Source-level ncu capture of the LJ column traversal through the torch-free A/B program (scripts/lj_col_ab.cu ->
# scripts/_bin/lj_col_ab): one `--set full` capture of a timed launch with SASS/source attribution, exported as CSV
# pages that can be read without a GPU (the library is built with -lineinfo).
# Run on the GPU box:   gpurun -- 'bash scripts/profile_lj.sh r02'
# Outputs (gpurun_out/): <tag>_lj.ncu-rep, <tag>_lj_raw.csv, <tag>_lj_source.csv, <tag>_lj_ab.json
tag=${1:-r02}
pat=${2:-ljColumnTraversal}
mkdir -p gpurun_out
scripts/_bin/lj_col_ab > gpurun_out/${tag}_lj_ab.json
ncu --set full --import-source on --clock-control none -k regex:${pat} -s 2 -c 1 -f -o gpurun_out/${tag}_lj scripts/_bin/lj_col_ab 63 2 > gpurun_out/${tag}_lj_ncu.log 2>&1
ncu -i gpurun_out/${tag}_lj.ncu-rep --page raw --csv > gpurun_out/${tag}_lj_raw.csv
ncu -i gpurun_out/${tag}_lj.ncu-rep --page source --csv > gpurun_out/${tag}_lj_source.csv 2>/dev/null || true
cat gpurun_out/${tag}_lj_ab.json
