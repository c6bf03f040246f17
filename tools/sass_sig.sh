#!/bin/bash
# Codegen signature of the four hot K1 kernels of a built library: instruction count, spill instructions and a
# hash of the opcode sequence.  Two builds of the same source are the same code iff the hashes agree (they do for
# the single-module build; `--split-compile` builds differ from run to run, see aas_enhancement_b200/build.py).
# usage: tools/sass_sig.sh [aas_enhancement_b200/libaas_lmfb.so]
LIB=${1:-aas_enhancement_b200/libaas_lmfb.so}
TMP=$(mktemp)
for fn in lmfb_k1ILi1ELb0ELi5ELi3ELb0ELb0E lmfb_k1ILi1ELb1ELi5ELi3ELb0ELb0E lmfb_k1ILi1ELb0ELi8ELi2ELb0ELb0E lmfb_k1ILi1ELb1ELi8ELi2ELb0ELb0E; do
  cuobjdump -sass "$LIB" | awk -v fn=$fn '/Function : /{p=index($0,fn)>0} p' | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" \
    | sed 's#/\* 0x[0-9a-f]* \*/##; s/[ \t]*$//' | awk '{ $1=""; print }' > "$TMP"
  echo "$fn instructions=$(wc -l < "$TMP") STL=$(grep -c STL "$TMP") LDL=$(grep -c LDL "$TMP") opcodes=$(awk '{print substr($1,1,1)=="@"?$2:$1}' "$TMP" | md5sum | cut -c1-8)"
done
rm -f "$TMP"
