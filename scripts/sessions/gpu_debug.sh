#!/bin/bash
echo "== default (mma) dbg tile 20"; SFB_MMA_DBG=20 timeout 100 python -u scripts/debug_grads.py > gpurun_out/dbg.log 2>&1; echo rc=$?; tail -48 gpurun_out/dbg.log
