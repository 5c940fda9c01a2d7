#!/bin/bash
out=gpurun_out/lstm_probe.log
: > $out
run() { echo "=== $*" >> $out; timeout 180 python tools/lstm_probe.py "$@" >> $out 2>&1; echo "rc=$?" >> $out; }
run --ts 0 --tile 16 --B 40 --T 6
run --ts 1 --tile 32 --B 70 --T 9
run --ts 1 --tile 16 --B 256 --T 300 --time --prof
run --ts 0 --tile 16 --B 256 --T 300 --time --prof
run --ts 1 --tile 32 --B 256 --T 300 --time --prof
run --ts 1 --tile 32 --B 1024 --T 300 --time
grep -vE "^  |Traceback" $out
