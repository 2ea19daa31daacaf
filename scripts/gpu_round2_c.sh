#!/bin/bash
for W in 16 24 32; do echo "== RG_WARPS=$W"; SPLPAK_B200_RG_WARPS=$W timeout 300 python scripts/eval_ab.py 1e9 3 3 2>&1 | grep regroup; done
