#!/bin/bash
# one GPU session: tests -> bench -> ncu launch lists (each ncu run only after its plain run exited 0)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
python bench.py --steps 30 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/bench.log
python tools/profile_step.py train 256 3 > gpurun_out/plain_train.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train.csv python tools/profile_step.py train 256 3 > gpurun_out/ncu_train.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_train.csv | head -40
