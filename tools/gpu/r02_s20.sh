#!/bin/bash
# round-2 GPU session 20: launch lists (ncu, per-launch durations) of the DGCNN forward and of the SSG step
mkdir -p gpurun_out
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_dgcnn_launches.csv python tools/profile_dgcnn.py 2 > gpurun_out/r02_dgcnn_prof.log 2>&1
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_step_launches.csv python tools/profile_step.py 2 > gpurun_out/r02_step_prof.log 2>&1
tail -2 gpurun_out/r02_dgcnn_prof.log gpurun_out/r02_step_prof.log
