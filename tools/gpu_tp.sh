#!/bin/bash
TP_DETAIL=bn_ timeout 300 python tools/train_profile.py 2>&1 | grep " us " | head -130 | awk '{printf "%s:%s ", substr($3,1,40), $1} END {print ""}'
