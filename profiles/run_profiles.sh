#!/bin/bash
# Round-1 profiling pass (run on the GPU box through gpurun; outputs land in gpurun_out/ and the
# summaries are copied into profiles/ by hand).  Numbers printed by runs under ncu are not bench values.
set -x
TAG=${1:-r1}
# 1. launch list of the bench command (per-launch durations, cold-cache and serialised)
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
# 2. full capture of the dominant kernel inside the C1 run (N = 2^10, one bold step in the middle of the run)
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:scalar_step_kernel -s 120 -c 1 \
    -o gpurun_out/ncu_${TAG}_c1_step -f python profiles/prof_c1.py 200 1024 1 > gpurun_out/ncu_${TAG}_c1.log 2>&1
# 3. the same kernel at a sample count that fills the machine (N = 2^17, orders 0:4) and at orders 0:6
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:scalar_step_kernel -s 6 -c 1 \
    -o gpurun_out/ncu_${TAG}_o4_bigN -f python profiles/throughput.py 4 131072 > gpurun_out/ncu_${TAG}_o4.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:scalar_step_kernel -s 6 -c 1 \
    -o gpurun_out/ncu_${TAG}_o6 -f python profiles/throughput.py 6 16384 > gpurun_out/ncu_${TAG}_o6.log 2>&1
# 4. sector-block models: full capture of the block walker (two-band model, orders 0:3, N = 2^12), its throughput
#    vs N, and the Hubbard-dimer (C3) / two-band (C4) configurations as whole inchworm! runs next to the CPU port
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:block_walk_kernel -s 3 -c 1 \
    -o gpurun_out/ncu_${TAG}_block_walk -f python profiles/throughput_block.py 3 4096 > gpurun_out/ncu_${TAG}_bw.log 2>&1
python profiles/throughput_block.py 3 1024 4096 16384 > gpurun_out/${TAG}_throughput_block.log 2>&1
python profiles/bench_c34.py c3 3 64 32768 8 2>/dev/null | tail -1 > gpurun_out/${TAG}_bench_c3.json
python profiles/bench_c34.py c4 3 32 1024 2 2>/dev/null | tail -1 > gpurun_out/${TAG}_bench_c4.json
python profiles/bench_c34.py c4 4 6 256 0 2>/dev/null | tail -1 > gpurun_out/${TAG}_bench_c4_orders04.json
# 5. condense the reports here (the .ncu-rep files with imported sources exceed what gpurun copies back)
LIB=qinchworm.jl_b200/libqinchworm_cuda.so
for r in c1_step o4_bigN o6; do
    python profiles/ncu_summary.py gpurun_out/ncu_${TAG}_${r}.ncu-rep > gpurun_out/${TAG}_ncu_${r}_summary.csv
    python profiles/ncu_lines.py gpurun_out/ncu_${TAG}_${r}.ncu-rep scalar_step_kernel $LIB 30 > gpurun_out/${TAG}_ncu_${r}_lines.txt
done
python profiles/ncu_summary.py gpurun_out/ncu_${TAG}_block_walk.ncu-rep > gpurun_out/${TAG}_ncu_block_walk_summary.csv
python profiles/ncu_lines.py gpurun_out/ncu_${TAG}_block_walk.ncu-rep block_walk $LIB 30 > gpurun_out/${TAG}_ncu_block_walk_lines.txt
rm -f gpurun_out/ncu_${TAG}_*.ncu-rep
