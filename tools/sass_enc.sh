#!/bin/bash
# dev helper: compile smaq_pack.cu (default widths only) and print static SASS statistics of the hot encode kernel
cd /root/repo/smart-quantization_b200/csrc || exit 1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -prec-div=true -prec-sqrt=true -ftz=false -fmad=false -DSMAQ_PACK_MINIMAL $SMAQ_EXTRA -Xptxas -v -c smaq_pack.cu -o /tmp/smaq_pack.o 2>&1 | grep -E "error|Used|spill" | head -${1:-6}
K=$(cuobjdump -sass /tmp/smaq_pack.o | grep -o "_ZN4smaq13encode_kernelILi5ELi2ELb1ELb0ELb0E[A-Za-z0-9_]*" | head -1)
cuobjdump -sass -fun "$K" /tmp/smaq_pack.o | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed -E 's/^\s+\/\*([0-9a-f]+)\*\/\s+//; s/\s*\/\*.*$//' > /tmp/enc.sass
wc -l /tmp/enc.sass
grep -n "LDS.128\|SHFL.UP PT, R[0-9]*, R[0-9]*, 0x1,\|BAR.SYNC" /tmp/enc.sass | head -12
