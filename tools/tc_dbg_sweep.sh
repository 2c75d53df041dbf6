#!/bin/bash
# Engine isolation of the strip kernel (results are garbage with DBG != 0): which role bounds a tile?
# DBG bits: 1 no MMAs, 2 no strip loads, 4 no conversion, 8 no output stores, 16 chains not read by the flush warps
for d in 0 1 2 4 8 16 6 24 30 31; do SGPU_FIR_TC_DBG=$d python tools/tc_quick.py 27 ${@:-512 256}; done 2>&1 | grep -v "^$"
