#!/bin/bash
# Final round-2 evidence on one B200 (outputs under gpurun_out/, summarised into profiles/ afterwards).
# Order: ncu capture of the dominant kernel -> traffic stamp -> bench lines (so that the line carries roofline.traffic).
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_lm -s 2 -c 1 -o gpurun_out/r2_lm_A7 -f python tools/dev_chain.py --check 0 --iters 1 > gpurun_out/r2_ncu_lm.log 2>&1; echo "ncu lm rc=$?"
python tools/stamp_traffic.py gpurun_out/r2_lm_A7.ncu-rep "dram__bytes_read.sum + dram__bytes_write.sum of the ONE kernel of A^7 (k_lm: evaluated as A x A^6), ncu --set full capture profiles/r2_lm_A7_details.txt; algorithmic bytes 221545092" > gpurun_out/r2_stamp.log 2>&1; echo "stamp rc=$?"
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_n1.json 2> /dev/null; echo "reference rc=$?"
python bench.py --bits 32 --no-cpu-baseline > gpurun_out/r2_bench_n1_u32.json 2> /dev/null; echo "u32 rc=$?"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu_bench.log 2>&1; echo "ncu list rc=$?"
sparse_linear_algebra_tests_b200/csrc/b200_bench --config sweep > gpurun_out/r2_sweep.csv 2>/dev/null; echo "sweep rc=$?"
