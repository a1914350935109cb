#!/bin/bash
# static instruction mix of one kernel: tools/sass_stats.sh <objdir> <group> <mangled-substring>
obj=$(realpath $1)/kernels_g$2.o
tmp=$(mktemp -d)
(cd $tmp && cuobjdump -xelf all $obj >/dev/null && nvdisasm *.cubin > all.sass)
awk -v pat="$3" '/^\.text\./{p=index($0,pat)>0} p' $tmp/all.sass | grep -E "^\s*/\*[0-9a-f]{4}\*/" | sed 's/\/\* 0x.*//' > $tmp/k.txt
echo "total $(wc -l < $tmp/k.txt)  moves $(grep -cE 'IMAD\.MOV|[^I]MOV |CS2R|HFMA2' $tmp/k.txt)  bssy $(grep -c BSSY $tmp/k.txt)  bra $(grep -c 'BRA' $tmp/k.txt)  sel $(grep -cE ' SEL |FSEL' $tmp/k.txt)"
rm -rf $tmp
