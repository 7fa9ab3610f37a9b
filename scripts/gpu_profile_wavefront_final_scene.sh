mkdir -p gpurun_out
VECCHIO_WF_SLOTS=1048576 python scripts/render_once.py final_scene 64 2 > gpurun_out/plain_wf.log 2>&1 && \
VECCHIO_WF_SLOTS=1048576 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_issued.avg.pct_of_peak_sustained_active --clock-control none -s 60 -c 90 --csv --log-file gpurun_out/launches_wf_fs.csv \
    python scripts/render_once.py final_scene 64 2 > gpurun_out/ncu_wf_launch.log 2>&1; echo "launch list rc=$?"
cat gpurun_out/plain_wf.log
