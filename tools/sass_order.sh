#!/bin/bash
# Prints the sequence of load / math groups of one kernel's SASS (how early the loads are issued).
# usage: tools/sass_order.sh <lib.so> <mangled-substring>
TMP=$(mktemp -d); cd $TMP
cuobjdump -xelf all "$1" >/dev/null
nvdisasm -c *.cubin 2>/dev/null | sed -n "/.text.*$2/,/^\s*\.section/p" | grep -n "LDG\|LDGSTS\|BAR\|MUFU\|DADD\|DMUL\|STS\|STG\|LDS\|CCTL" | sed 's/\/\*[0-9a-f]*\*\///' | awk '{n=$1; sub(":","",n); op=$2; if (op ~ /^@/) op=$3; split(op,a,"."); o=a[1]; if (o!=prev) { if (prev!="") print start"-"last, prev, cnt; start=n; cnt=0; prev=o } cnt++; last=n } END {print start"-"last, prev, cnt}'
