#!/bin/bash
# usage: scripts/spill_lines.sh <object.o> <kernel-name-substring>   -- source lines of the local-memory (spill) instructions
set -e
T=$(mktemp -d); cd $T
cuobjdump -xelf all "$1" > /dev/null
nvdisasm -g -c *.cubin > all.dis 2>/dev/null
awk -v k="$2" '/^\/\/-+ \.text\./{on = index($0, k) > 0} on && /## File/{line=$0} on && /STL|LDL/{print line}' all.dis \
  | sed 's/.*line \([0-9]*\).*/\1/' | sort -n | uniq -c
rm -rf $T
