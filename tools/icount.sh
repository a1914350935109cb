#!/bin/bash
# warp/thread instruction counts of the photon kernel for one deck (ncu, 2 cheap metrics): tools/icount.sh cube60b:1e6
ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:photon_kernel -c 1 \
    python tools/perf_sweep.py "$@" 2>&1 | grep -E "smsp__|gpu__time|photon_kernel<" | sed 's/  */ /g'
