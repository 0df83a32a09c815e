#!/bin/bash
# Last GPU seconds of round 2: the C++ drop-in demo (no Python start-up) in KM mode on the committed 700 x 640 scene,
# three times: default route, every settled iteration's candidate block forced to overflow, general route only.
mkdir -p gpurun_out/c17 && cd gpurun_out/c17
D=../../gh-icp_b200/cxx/dropin_demo
S=../../tools/scenes/dropin_km_700x640.bin
timeout 8 $D $S 60 > default.txt 2>&1; echo "rc=$?" >> default.txt
GHICP_KM_XUSE_MAX=8 timeout 8 $D $S 60 > overflow.txt 2>&1; echo "rc=$?" >> overflow.txt
GHICP_KM_GENERAL=1 timeout 8 $D $S 60 > general.txt 2>&1; echo "rc=$?" >> general.txt
cmp default.txt overflow.txt && cmp default.txt general.txt && echo IDENTICAL
