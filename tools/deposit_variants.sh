#!/bin/bash
# profiles/r2_deposit_variants.md: kernel time + L2 reduction sectors of the deposit variants, and the reduction-placement probe
out=gpurun_out/r2_deposit_variants.log
: > $out
M=gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_red.sum,lts__t_sectors_srcunit_tex_op_red_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_red_lookup_miss.sum,dram__bytes_write.sum,smsp__inst_executed.sum
for v in default matchany smemtile; do
  if [ $v = default ]; then unset MCXB200_LIB; else export MCXB200_LIB=$PWD/mcxcl_b200/build/variants/$v/libmcxb200.so; fi
  echo "== $v (kernel-window timing, 1e8 / 3e7 photons)" >> $out
  timeout 600 python tools/perf_sweep.py cube60:1e8 cube60b:1e8 colin27:3e7 >> $out 2>&1
  for deck in cube60 cube60b colin27; do
    echo "== $v ncu $deck 1e7" >> $out
    timeout 600 ncu --metrics $M --clock-control none -k regex:photon_kernel -s 1 -c 1 --csv python tools/ncu_one.py $deck 1e7 2>&1 | grep -E "photon_kernel" | awk -F'","' '{print $(NF-2), $(NF)}' >> $out
  done
done
unset MCXB200_LIB
for mode in 0 1 2 3 4; do
  echo "== red_die mode $mode" >> $out
  timeout 300 ncu --metrics lts__t_sectors_srcunit_tex_op_red.sum,lts__t_sectors_srcunit_tex_op_red_lookup_hit.sum,lts__t_sectors_srcunit_tex_op_red_lookup_miss.sum,gpu__time_duration.sum --clock-control none -k regex:hook_red_die --csv python tools/red_die.py $mode 2>&1 | grep -E "hook_red_die|issued" | awk -F'","' '{ if (NF>3) print $(NF-2), $(NF); else print $0 }' >> $out
done
cat $out
