#!/bin/bash
# Round-2 profiling pass (run on the GPU box through gpurun; outputs land in gpurun_out/ and the summaries are copied
# into profiles/ by hand).  Numbers printed by runs under ncu are not bench values.
set -x
TAG=${1:-r2}
LIB=qinchworm.jl_b200/libqinchworm_cuda.so
# 1. launch list of the bench command (per-launch durations, cold-cache and serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-stress --no-extra > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
# 2. the persistent run kernel on the README configuration (one launch = all 198 bold steps)
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:scalar_run_kernel -s 2 -c 1 \
    -o gpurun_out/ncu_${TAG}_c1_run -f python profiles/prof_c1.py 200 1024 4 > gpurun_out/ncu_${TAG}_c1_run.log 2>&1
# 3. the step kernel with the machine full: orders 0:4 at N = 2^17, orders 0:6 at N = 2^14
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:scalar_step_kernel -s 6 -c 1 \
    -o gpurun_out/ncu_${TAG}_o4_bigN -f python profiles/throughput.py 4 131072 > gpurun_out/ncu_${TAG}_o4.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:scalar_step_kernel -s 6 -c 1 \
    -o gpurun_out/ncu_${TAG}_o6 -f python profiles/throughput.py 6 16384 > gpurun_out/ncu_${TAG}_o6.log 2>&1
# 4. sector-block models: the shipped walker (one 24-warp CTA per SM; two-band model, orders 0:3, N = 2^12) and the
#    FP64 tensor-core kernel for blocks of 5 to 8 rows next to the general FMA kernel
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:block_walk_kernel -s 3 -c 1 \
    -o gpurun_out/ncu_${TAG}_block_walk -f python profiles/throughput_block.py 3 4096 > gpurun_out/ncu_${TAG}_bw.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:block_mma_kernel -s 3 -c 1 \
    -o gpurun_out/ncu_${TAG}_block_mma -f python profiles/throughput_mma.py 2 4096 > gpurun_out/ncu_${TAG}_mma.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:block_step_kernel -s 3 -c 1 \
    -o gpurun_out/ncu_${TAG}_block_fma8 -f python profiles/throughput_mma.py 2 4096 > gpurun_out/ncu_${TAG}_fma8.log 2>&1
python profiles/throughput_mma.py 2 256 1024 4096 > gpurun_out/${TAG}_throughput_mma.log 2>&1
python profiles/throughput_block.py 3 1024 4096 16384 > gpurun_out/${TAG}_throughput_block.log 2>&1
python profiles/throughput.py 4 1024 16384 131072 1048576 > gpurun_out/${TAG}_throughput_o4.log 2>&1
python profiles/throughput.py 6 16384 > gpurun_out/${TAG}_throughput_o6.log 2>&1
python profiles/bench_c34.py c4 3 32 1024 2 2>/dev/null | tail -1 > gpurun_out/${TAG}_bench_c4.json
python profiles/bench_c34.py c3 4 64 4096 1 2>/dev/null | tail -1 > gpurun_out/${TAG}_bench_c3.json
python profiles/bench_c2.py 128 1024 8192 65536 > gpurun_out/${TAG}_bench_c2.log 2>&1
python profiles/run_vs_step.py 200 1024 2048 4096 > gpurun_out/${TAG}_run_vs_step.log 2>&1
# 5. condense the reports here (the .ncu-rep files with imported sources exceed what gpurun copies back)
python profiles/ncu_summary.py gpurun_out/ncu_${TAG}_c1_run.ncu-rep > gpurun_out/${TAG}_ncu_c1_run_summary.csv
python profiles/ncu_lines.py gpurun_out/ncu_${TAG}_c1_run.ncu-rep scalar_run_kernelILb1 $LIB 40 > gpurun_out/${TAG}_ncu_c1_run_lines.txt
for r in o4_bigN o6; do
    python profiles/ncu_summary.py gpurun_out/ncu_${TAG}_${r}.ncu-rep > gpurun_out/${TAG}_ncu_${r}_summary.csv
    python profiles/ncu_lines.py gpurun_out/ncu_${TAG}_${r}.ncu-rep scalar_step_kernelILb1 $LIB 30 > gpurun_out/${TAG}_ncu_${r}_lines.txt
done
python profiles/ncu_summary.py gpurun_out/ncu_${TAG}_block_walk.ncu-rep > gpurun_out/${TAG}_ncu_block_walk_summary.csv
python profiles/ncu_lines.py gpurun_out/ncu_${TAG}_block_walk.ncu-rep block_walk_kernel $LIB 30 > gpurun_out/${TAG}_ncu_block_walk_lines.txt
python profiles/ncu_summary.py gpurun_out/ncu_${TAG}_block_mma.ncu-rep > gpurun_out/${TAG}_ncu_block_mma_summary.csv
python profiles/ncu_lines.py gpurun_out/ncu_${TAG}_block_mma.ncu-rep block_mma_kernel $LIB 30 > gpurun_out/${TAG}_ncu_block_mma_lines.txt
python profiles/ncu_summary.py gpurun_out/ncu_${TAG}_block_fma8.ncu-rep > gpurun_out/${TAG}_ncu_block_fma8_summary.csv
rm -f gpurun_out/ncu_${TAG}_*.ncu-rep
