#!/bin/bash
# usage: sass_stats.sh <mangled-substring>   -> opcode histogram of that kernel in libtfb_b200.so
pat="$1"
cuobjdump -sass topoflow36_b200/libtfb_b200.so | awk -v pat="$pat" '/Function : /{f=($0 ~ pat)} f' > /tmp/k.sass
grep -E "^\s+/\*[0-9a-f]{4}\*/" /tmp/k.sass | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//; s/^@!?U?P[0-9T]+ //' | awk '{print $1}' | sed 's/\..*//; s/;//' | sort | uniq -c | sort -rn | awk '{printf "%s:%s ", $2, $1} END{print ""}'
echo "total: $(grep -cE "^\s+/\*[0-9a-f]{4}\*/" /tmp/k.sass)"
