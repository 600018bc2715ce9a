#!/bin/bash
# experiment: two candidates per trip of a row walk (library variant built with -DS2B_ROW_WIDTH=2)
W2=$PWD/srrg2_slam_interfaces_b200/libsrrg2b_w2.so
SRRG2B_LIB=$W2 python tools/iter_profile.py 1000000 4
python tools/iter_profile.py 1000000 4
SRRG2B_LIB=$W2 timeout 20 python -m pytest tests/test_gpu_parity_icp.py -m gpu -x -q 2>&1 | tail -2
