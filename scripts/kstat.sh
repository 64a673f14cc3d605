#!/bin/bash
# Static report of the trace kernels of a built library: registers / spills (recompiled with -Xptxas -v) and
# SASS size of trace_kernel<1,0>.   scripts/kstat.sh [extra nvcc flags]
cd "$(dirname "$0")/../pyrayt_b200/csrc"
nvcc -O3 -std=c++17 -lineinfo -fmad=false -gencode arch=compute_100a,code=sm_100a "$@" -Xptxas -v -c -o /tmp/kstat.o prt_kernels.cu 2>&1 \
 | awk '/Compiling entry function/{name=$0; sub(/.*function ./,"",name); sub(/. for.*/,"",name)} /spill/{sp=$0} /Used [0-9]+ registers/{ if (name ~ /trace_kernel/) print name, "|", sp, "|", $0 }' | sed 's/ptxas info    ://g; s/bytes stack frame/stack/; s/bytes spill stores/sp.st/; s/bytes spill loads/sp.ld/' | cut -c1-200
cuobjdump -sass /tmp/kstat.o 2>/dev/null | awk '/Function :/{name=$3} /^ +\/\*[0-9a-f]{4}\*\//{n[name]++} END{for(k in n) if (k ~ /trace_kernel/) print n[k], "SASS instr", k}' | sort -n
