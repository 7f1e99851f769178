#!/bin/bash
# Round-2 evidence on one B200 (outputs under gpurun_out/, summarised into profiles/ afterwards)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; tail -2 gpurun_out/r2_pytest.log
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_n1.json 2> /dev/null; echo "reference rc=$?"
python bench.py --bits 32 --no-cpu-baseline > gpurun_out/r2_bench_n1_u32.json 2> /dev/null
python tools/big_config.py rmat --scale 20 --check 1 --iters 3 > gpurun_out/r2_big.txt 2>&1
python tools/big_config.py rmat --scale 22 --check 0 --iters 3 >> gpurun_out/r2_big.txt 2>&1
python tools/big_config.py rmat --scale 17 --abc 0.57 0.19 0.19 --check 1 --iters 3 >> gpurun_out/r2_big.txt 2>&1
python tools/big_config.py rmat --scale 20 --abc 0.57 0.19 0.19 --check 0 --iters 2 >> gpurun_out/r2_big.txt 2>&1
python tools/big_config.py torus --side 100 --power 5 --check 1 --iters 3 --device-build 1 >> gpurun_out/r2_big.txt 2>&1
grep -E "^A\^|rmat:|torus" gpurun_out/r2_big.txt | cut -c1-230
sparse_linear_algebra_tests_b200/csrc/b200_bench --config sweep > gpurun_out/r2_sweep.csv 2>/dev/null
# ncu: launch list of a short bench, then full captures of the dominant kernels (numbers printed under ncu are not bench values)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_ncu_bench.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k k_rw_fused -s 3 -c 1 -o gpurun_out/r2_rwfused_A7 -f python tools/dev_chain.py --check 0 --iters 1 > gpurun_out/r2_ncu_rw.log 2>&1; echo "ncu rw rc=$?"
ncu --set full --clock-control none --import-source on -k k_hv -c 2 -o gpurun_out/r2_hv_g500 -f python tools/big_config.py rmat --scale 17 --abc 0.57 0.19 0.19 --check 0 --iters 1 > gpurun_out/r2_ncu_hv.log 2>&1; echo "ncu hv rc=$?"
ncu --set full --clock-control none --import-source on -k k_num_cta -s 2 -c 1 -o gpurun_out/r2_numcta_rmat18 -f python tools/big_config.py rmat --scale 18 --check 0 --iters 1 > gpurun_out/r2_ncu_cta.log 2>&1; echo "ncu cta rc=$?"
