#!/bin/bash
mkdir -p gpurun_out
python tools/pcie_probe.py > gpurun_out/r02_pcie_probe.txt 2>&1; cat gpurun_out/r02_pcie_probe.txt
