#!/bin/bash
# decomposition of the 16-bit GEMM's time: full kernel / without the TMA stores / TMEM reads only
for d in 0 1 3; do echo "DEER_H16_DEBUG=$d"; DEER_H16_DEBUG=$d timeout 100 python tools/gemm_h16_probe.py 2>&1 | grep "h16" | grep -v block0 | cut -c1-75; done
