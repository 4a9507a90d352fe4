#!/bin/bash
RS="6 8 10" YS="20" bash scripts/sweep_queue2.sh "-DCCU_Q_UNROLL2;-DCCU_Q_WARPS=28"
