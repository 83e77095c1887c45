python tools/pcie_probe.py > gpurun_out/pcie_probe.txt 2>&1
RB200_TRACE=1 python tools/e2e_breakdown.py > gpurun_out/e2e_numa.txt 2>&1
RB200_NO_NUMA=1 RB200_TRACE=1 python tools/e2e_breakdown.py > gpurun_out/e2e_nonuma.txt 2>&1
cat gpurun_out/pcie_probe.txt; grep -v "^rb200 foreach" gpurun_out/e2e_numa.txt | tail -10;  grep -v "^rb200 foreach" gpurun_out/e2e_nonuma.txt | tail -2
