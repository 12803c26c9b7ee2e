python tools/stage_under_load.py 1 2>&1 | tail -12
python tools/stage_under_load.py 8 2>&1 | tail -12
