#!/bin/bash
EXP_ONLY=1 timeout 600 python tools/exp_cycles.py 34 20 2>&1 | tail -8
