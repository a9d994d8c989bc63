timeout 600 python tools/prof_joint.py 20000 30 2>&1 | tail -6
