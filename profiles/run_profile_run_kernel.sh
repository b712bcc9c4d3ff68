#!/bin/bash
# ncu capture of the persistent run kernel on the README configuration (C1): one launch = all 198 bold steps.
TAG=${1:-r2}
LIB=qinchworm.jl_b200/libqinchworm_cuda.so
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:scalar_run_kernel -s 2 -c 1 \
    -o gpurun_out/ncu_${TAG}_c1_run -f python profiles/prof_c1.py 200 1024 4 > gpurun_out/ncu_${TAG}_c1_run.log 2>&1
python profiles/ncu_summary.py gpurun_out/ncu_${TAG}_c1_run.ncu-rep > gpurun_out/${TAG}_ncu_c1_run_summary.csv
python profiles/ncu_lines.py gpurun_out/ncu_${TAG}_c1_run.ncu-rep scalar_run_kernelILb1 $LIB 60 > gpurun_out/${TAG}_ncu_c1_run_lines.txt
rm -f gpurun_out/ncu_${TAG}_c1_run.ncu-rep
