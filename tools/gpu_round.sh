#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
out=gpurun_out/lstm_probe.log; : > $out
run() { echo "=== $*" >> $out; timeout 180 python tools/lstm_probe.py "$@" >> $out 2>&1; echo "rc=$?" >> $out; }
run --ts 1 --tile 16 --B 256 --T 300 --time --prof
grep -E "prof|time" $out
python bench.py --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 2800 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
python tools/profile_step.py train 256 3 > gpurun_out/plain_train.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python tools/profile_step.py train 256 3 > gpurun_out/ncu_train.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_train.csv | head -24
