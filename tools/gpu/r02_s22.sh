#!/bin/bash
# round-2 GPU session 22: profile artefacts of the current build -- launch list of the bench command, ncu --set full of the
# dominant kernels (dual-group SA levels, F1, kNN), traffic numbers, and the error statistics of the kNN bound
mkdir -p gpurun_out
timeout -k 10 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-extra --no-train --no-cpu-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:tc_sa_dual_kernel --launch-skip 4 -c 2 -o gpurun_out/r02_dual_full -f python tools/profile_step.py 3 > gpurun_out/r02_ncu_dual.log 2>&1
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:tc_dense3_kernel --launch-skip 14 -c 7 -o gpurun_out/r02_dense3_full -f python tools/profile_step.py 3 > gpurun_out/r02_ncu_dense3.log 2>&1
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel --launch-skip 2 -c 1 -o gpurun_out/r02_knn_full10 -f python tools/knn_tc_timing.py > gpurun_out/r02_ncu_knn10.log 2>&1
PSA_LIB_PATH=$PWD/scanobjectnn_b200/libpsa_errstat.so timeout -k 10 300 python tools/knn_tc_timing.py > gpurun_out/r02_knn_errstat3.json 2>gpurun_out/r02_knn_errstat3.err
cat gpurun_out/r02_knn_errstat3.json
timeout -k 10 900 python tools/profile_train.py > gpurun_out/r02_train_profile.json 2> gpurun_out/r02_train_profile.err; tail -c 600 gpurun_out/r02_train_profile.json
ls -la gpurun_out/*.ncu-rep | tail -5
