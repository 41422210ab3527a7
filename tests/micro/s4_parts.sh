#!/bin/bash
# timing experiment: which part of a tick of sweep4_kernel costs what (results are garbage with any bit set)
export LDU_STENCIL=4 LDU_STENCIL_M=2
for d in 0 1 2 4 8 16 32 64 3 7 15 31 63 127; do
  LDU_S3_DBG=$d LDU_S2_TRACE=/tmp/tr.txt python tests/perf_sweeps.py 216 32 4 3 > /dev/null 2>&1
  awk -v d=$d '{printf "dbg %3d q %d cycles/tick %d\n", d, $4, $11/250}' /tmp/tr.txt
done
