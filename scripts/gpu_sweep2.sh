#!/bin/bash
mkdir -p gpurun_out
T="96:2048:0:3:12:100:30"
U="96:512:150:2:4:100:30"
S="1:$T:1024:1024,1:$T:256:256,1:$T:128:128,1:$T:64:64,1:$T:32:32,1:$T:64:1024,1:$T:1024:64,1:$T:32:128,1:$T:128:32"
S="$S,6:$U:1024:1024,6:$U:256:256,6:$U:128:128,6:$U:64:64,6:$U:32:32,6:$U:64:1024,6:$U:1024:64,9:$U:64:64,9:$U:128:128,0:$T:64:64,5:$T:64:64"
timeout 600 python scripts/sweep_em.py 300 "$S" > gpurun_out/sweep2.txt 2>&1
SB_EM_GROUP_CM=64 SB_EM_GROUP_TM=64 timeout 200 python scripts/timeline_em.py 500000 > gpurun_out/timeline_g64.txt 2>&1
SB_EM_CONFIG=6 SB_EM_LWARP=512 SB_EM_BALANCE=150 SB_EM_GROUP_CM=64 SB_EM_GROUP_TM=64 timeout 200 python scripts/timeline_em.py 500000 > gpurun_out/timeline_cfg6_g64.txt 2>&1
cat gpurun_out/sweep2.txt
