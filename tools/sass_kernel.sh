#!/bin/bash
# usage: tools/sass_kernel.sh <group> <mangled-substring> > out.sass   (nvdisasm listing with line info of one kernel)
set -e
obj=/root/repo/mcxcl_b200/build/kernels_g$1.o
tmp=$(mktemp -d)
(cd $tmp && cuobjdump -xelf all $obj >/dev/null && nvdisasm --print-line-info *.cubin > all.sass)
awk -v pat="$2" '/^\.text\./{p=index($0,pat)>0} p' $tmp/all.sass
rm -rf $tmp
