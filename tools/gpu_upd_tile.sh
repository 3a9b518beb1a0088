#!/bin/bash
for u in 64 32 16; do echo "== upd_tile $u"; GPUHASH_UPD_TILE=$u EXP_ONLY=1 timeout 600 python tools/exp_cycles.py 34 20 2>&1 | grep '"mixed"' ; done
