#!/bin/bash
# Profiling pass of the scalar step kernel (run on the GPU box through gpurun; summaries land in gpurun_out/).
# Numbers printed by runs under ncu are not bench values.
TAG=${1:-r2}
LIB=qinchworm.jl_b200/libqinchworm_cuda.so
K=${2:-scalar_step_kernel}
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$K -s 6 -c 1 \
    -o gpurun_out/ncu_${TAG}_o4_bigN -f python profiles/throughput.py 4 131072 > gpurun_out/ncu_${TAG}_o4.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$K -s 6 -c 1 \
    -o gpurun_out/ncu_${TAG}_o6 -f python profiles/throughput.py 6 16384 > gpurun_out/ncu_${TAG}_o6.log 2>&1
for r in o4_bigN o6; do
    python profiles/ncu_summary.py gpurun_out/ncu_${TAG}_${r}.ncu-rep > gpurun_out/${TAG}_ncu_${r}_summary.csv
    python profiles/ncu_lines.py gpurun_out/ncu_${TAG}_${r}.ncu-rep ${KSYM:-scalar_step_kernelILb1} $LIB 40 > gpurun_out/${TAG}_ncu_${r}_lines.txt
done
rm -f gpurun_out/ncu_${TAG}_*.ncu-rep
