#!/bin/bash
mkdir -p gpurun_out
SPLPAK_B200_PANELCLK=1 timeout 300 python bench.py --steps 3 --warmup 1 --no-e2e --no-cpu > gpurun_out/r4j_clk.json 2> gpurun_out/r4j_clk.err; echo "clk rc=$?"; grep -E "data-flow|L10 product|communication|panel worker|helper 0|B1 arrivals" gpurun_out/r4j_clk.err | tail -6
