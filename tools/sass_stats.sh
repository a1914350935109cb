#!/bin/bash
# Static facts about the shipped kernels, from the build logs (ptxas -v) and the SASS of the object files:
#   tools/sass_stats.sh [objdir]      (default: mcxcl_b200/build)   > profiles/r2_sass_stats.txt
# For every listed kernel: registers, stack frame, spill bytes, and the static counts of the instructions the design
# rests on (REDG = fire-and-forget reductions, ATOMG = atomics with a return value, MUFU, BSSY, LDG, LDS/STS, CALL).
objdir=$(realpath ${1:-mcxcl_b200/build})
kernels=(
  "0 photon_kernelILi0ELb1ELi1EhdLb0ELb0ELi8E  pencil/reflect/det1/u8/f64/common/queue8   (cube60b headline)"
  "0 photon_kernelILi0ELb1ELi1EhdLb0ELb0ELi0E  pencil/reflect/det1/u8/f64/common/queue0   (colin27)"
  "0 photon_kernelILi0ELb0ELi1EhdLb0ELb0ELi8E  pencil/noreflect/det1/u8/f64/common/queue8 (cube60)"
  "1 photon_kernelILi8ELb1ELi0EhdLb0ELb0ELi8E  disk/reflect/det0/u8/f64/common/queue8     (skinvessel)"
  "2 photon_kernelILi6ELb1ELi0EhdLb0ELb0ELi0E  fourier/reflect/det0/u8/f64/common/queue0  (digimouse)"
  "5 photon_kernelILin1ELb1ELi1EhdLb0ELb1ELi0ELb0E any/reflect/det1/u8/f64/generic            (run-time options)"
  "8 photon_kernelILin1ELb1ELi1EhdLb0ELb1ELi0ELb1E any/reflect/det1/u8/f64/generic/ext        (polarised, RF, adjoint detector sources)"
  "9 photon_kernelILin1ELb1ELi1EjdLb0ELb1ELi0ELb1E any/reflect/det1/u32/f64/generic/ext       (split-voxel media)"
)
tmp=$(mktemp -d)
for k in "${kernels[@]}"; do
  set -- $k; g=$1; pat=$2; shift 2
  echo "== $* [$pat]"
  grep -A3 "Compiling entry function '_ZN4mcxb13${pat}" $objdir/kernels_g$g.log | grep -E "Used|spill" | sed 's/ptxas info    : //; s/^ *//'
  if [ ! -f $tmp/g$g.sass ]; then (cd $tmp && mkdir -p x$g && cd x$g && cuobjdump -xelf all $objdir/kernels_g$g.o >/dev/null && nvdisasm *.cubin > ../g$g.sass); fi
  awk -v pat="$pat" '/^\.text\./{p=index($0,pat)>0} p' $tmp/g$g.sass | grep -E "^\s*/\*[0-9a-f]{4,}\*/" > $tmp/k.txt
  printf "static SASS instructions %d:" $(wc -l < $tmp/k.txt)
  for op in "REDG.E.ADD.F64" "REDG.E.ADD.F32" ATOMG ATOMS MUFU BSSY BSYNC "BRA" CALL LDG LDS STS LDC LDCU S2R VOTE "IMAD.MOV\|[^I]MOV "; do
    printf "  %s %d" "$(echo $op | sed 's/\\|.*//')" $(grep -cE "$(echo $op | sed 's/\\|/|/')" $tmp/k.txt)
  done
  echo
done
rm -rf $tmp
