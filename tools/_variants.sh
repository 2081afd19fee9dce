python tools/ctl_timing.py 9472 1070000 2>&1 | grep "PhiX\|unassigned\|decode"
python -m pytest tests/test_gpu_parity.py -x -q -k control 2>&1 | tail -3
