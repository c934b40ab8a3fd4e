cd /root/repo; python tools/bp_trace.py 64 2>&1 | tail -4; python tools/bp_trace.py 1 2>&1 | tail -3
