timeout 400 python tools/prof_joint.py 5000 30 2>&1 | tail -4
