#!/bin/bash
# round-2 GPU session 4: training parity after the reduction / CSR / split-K rewrites, F1 v3b, MSG + pooling, ncu of F1, bench line
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_train_gpu.py -q -s > gpurun_out/r02_t4_train.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t4_train.log
timeout -k 10 900 python -m pytest tests/test_mlp_gpu.py -x -q -k "conv1_prebn or unit_scale or pooling or msg" > gpurun_out/r02_t4_misc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_t4_misc.log
timeout -k 10 120 python tools/f1_timing.py > gpurun_out/r02_f1v3b.json 2>gpurun_out/r02_f1v3b.err
PSA_LIB_PATH=$PWD/scanobjectnn_b200/libpsa_f1timing.so PSA_F1_TLOG=1 timeout -k 10 120 python tools/f1_timing.py > gpurun_out/r02_f1v3b_timeline.json 2>&1
timeout -k 10 300 ncu --set full --clock-control none --import-source on -k regex:sa_conv1_stream --launch-skip 2 -c 1 -o gpurun_out/r02_f1_full -f python tools/profile_ops.py > gpurun_out/r02_ncu_f1.log 2>&1
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_train_launches_b.csv python tools/profile_train.py 2 > gpurun_out/r02_train_prof_b.log 2>&1
timeout -k 10 900 python bench.py --steps 100 --warmup 5 --train-steps 10 --no-extra > gpurun_out/r02_bench_s4.json 2>gpurun_out/r02_bench_s4.err; echo "bench rc=$?" >> gpurun_out/r02_bench_s4.err
tail -12 gpurun_out/r02_t4_train.log; tail -4 gpurun_out/r02_t4_misc.log; cat gpurun_out/r02_f1v3b*.json; tail -3 gpurun_out/r02_bench_s4.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_s4.json"))
print(json.dumps({k: d[k] for k in ("value", "ms_per_step", "one_step_at_a_time", "train")}, indent=1))
print(json.dumps(d["roofline_f1"], indent=1))
PY
