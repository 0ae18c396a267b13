cd /root/repo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/stream_launches.csv python tools/stream_probe.py 56 > gpurun_out/stream_probe.log 2>&1
tail -2 gpurun_out/stream_probe.log
wc -l gpurun_out/stream_launches.csv
