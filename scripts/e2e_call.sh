#!/bin/bash
# gpurun call: host<->device copy rates and the traced phases of the end-to-end generator path
O=gpurun_out/e2e
mkdir -p $O
timeout 200 python scripts/pcie_probe.py > $O/pcie_probe.log 2>&1
cat $O/pcie_probe.log
WENDY_B200_TRACE=1 timeout 200 python scripts/e2e_phases.py > $O/e2e_phases_trace.log 2>&1
cat $O/e2e_phases_trace.log | grep -v "^Exception\|^Traceback\|File\|ImportError"
