#!/bin/bash
# final-build counters: ncu --set full of the headline kernel (cube60b 1e8) and of the digimouse kernel (3e7) after enter_volume was inlined
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:photon_kernel -s 1 -c 1 -f -o gpurun_out/r2_prof_cube60b_v11 python tools/ncu_one.py cube60b 1e8 > gpurun_out/r2_ncu_cube60b_v11.log 2>&1
tail -2 gpurun_out/r2_ncu_cube60b_v11.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:photon_kernel -s 1 -c 1 -f -o gpurun_out/r2_prof_digimouse_v11 python tools/ncu_one.py digimouse 3e7 > gpurun_out/r2_ncu_digimouse_v11.log 2>&1
tail -2 gpurun_out/r2_ncu_digimouse_v11.log
ls -la gpurun_out/*.ncu-rep
