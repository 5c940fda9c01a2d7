#!/bin/bash
# round-2 final GPU call: whole -m gpu suite, full bench line, launch lists, ncu --set full of the roofline kernels and of
# the LSTM kernels inside the real step, step timelines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_final.log | head -20
timeout 1200 python bench.py > gpurun_out/bench_final.log 2> gpurun_out/bench_final.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_final.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_final.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','launches_per_step')}, d['e2e']['ms_per_step']); print(d['inference']['ms_per_step'], d['inference']['e2e']['ms_per_step'], d['inference']['launches_per_step']); print(d.get('strong_scaling'))
    for r in d.get('pooled_model_sweep') or []: print(r)
    r=d['roofline']; print(r['frac'], r['us_per_launch'], r['nig_head_loss']['frac'], r['lstm_recurrence'], r['attn_pool']['frac']); print(d['config'].get('loss_check_vs_cpu_port')); print(d.get('cpu_baseline'))
except Exception as e:
    print("bench parse failed", e)
PY
for mode in train infer; do
  B=256; [ $mode = infer ] && B=1024
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$mode.csv python tools/profile_step.py $mode $B 3 > gpurun_out/ncu_$mode.log 2>&1
  python tools/summarize_launches.py gpurun_out/launches_$mode.csv > gpurun_out/launches_${mode}_summary.txt; head -12 gpurun_out/launches_${mode}_summary.txt
done
timeout 300 python tools/step_timeline.py > gpurun_out/timeline.log 2>&1; echo "timeline rc=$?"; tail -19 gpurun_out/timeline.log
timeout 300 python tools/step_timeline.py --input-grads > gpurun_out/timeline_ig.log 2>&1; echo "timeline-ig rc=$?"
# ncu --set full captures go to /tmp on the box (the reports are tens of MB); only the summaries travel back
timeout 400 ncu --set full --clock-control none -k "regex:gemm_h16|nig_loss" -o /tmp/r2_roofline python tools/ncu_roofline.py > gpurun_out/ncu_roofline.log 2>&1; echo "ncu roofline rc=$?"
python tools/ncu_summary.py /tmp/r2_roofline.ncu-rep > gpurun_out/r2_ncu_full_roofline_summary.txt 2>&1
timeout 400 ncu --set full --clock-control none -k "regex:lstm_(fwd|bwd)_cluster" -s 4 -c 4 -o /tmp/r2_lstm_step python tools/profile_step.py train 256 2 > gpurun_out/ncu_lstm.log 2>&1; echo "ncu lstm rc=$?"
python tools/ncu_summary.py /tmp/r2_lstm_step.ncu-rep > gpurun_out/r2_ncu_full_lstm_summary.txt 2>&1
timeout 400 ncu --set full --clock-control none -k "regex:attn_pool_fwd|lstm_fwd_cluster" -c 3 -o /tmp/r2_infer python tools/profile_step.py infer 1024 1 > gpurun_out/ncu_infer_full.log 2>&1; echo "ncu infer rc=$?"
python tools/ncu_summary.py /tmp/r2_infer.ncu-rep > gpurun_out/r2_ncu_full_infer_summary.txt 2>&1
grep -E "^==|gpu__time_duration|dram__bytes" gpurun_out/r2_ncu_full_lstm_summary.txt gpurun_out/r2_ncu_full_roofline_summary.txt gpurun_out/r2_ncu_full_infer_summary.txt | head -70
du -sh gpurun_out
