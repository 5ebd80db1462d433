timeout 300 python tests/gpu_diag.py conv_chain 2>&1 | cut -c1-200 | tail -7
timeout 60 python tools/conv_timers.py chain 2>&1 | grep layer | cut -c60-800
