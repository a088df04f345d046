#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -c "
import diffsinger_b200 as dsx
rc, txt = dsx.selftest(0, 3)
print(rc); print(txt)
" > gpurun_out/ingest.log 2>&1
cat gpurun_out/ingest.log
