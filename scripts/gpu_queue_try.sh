#!/bin/bash
export CCU_Q_MARCH_WARPS=64
RS="8" YS="8 20" bash scripts/sweep_queue2.sh "-DCCU_Q_ROLES;-DCCU_Q_ROLES -DCCU_Q_WARPS=32"
