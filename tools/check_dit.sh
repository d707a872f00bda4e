cd /root/repo
timeout 600 python tools/dit_leg.py n1 c4 > gpurun_out/dit_leg_new.jsonl 2> gpurun_out/dit_leg_new.err

