#!/bin/bash
# round-2 GPU session 42: F1 ring depth / batch size variants (instrumented libraries), bench with 8 streams
mkdir -p gpurun_out
for v in "" _ring5 _ring6 _b16 _b16r5; do
  echo "variant libpsa$v"
  PSA_LIB_PATH=scanobjectnn_b200/libpsa$v.so timeout -k 10 120 python tools/f1_timing.py 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['isolated_us_median'], d['steady_us'])"
done
