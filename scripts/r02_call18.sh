#!/bin/bash
export ACT_SCALE=0
echo default; python scripts/time_step.py
echo prio0; RR_SIDE_PRIORITY=0 python scripts/time_step.py
echo serial; RR_DEBUG_SERIAL_HEAVY=1 python scripts/time_step.py
echo heavy1-eager0; RR_EAGER_COLLIDE=0 python scripts/time_step.py
echo heavy0-eager0; RR_EAGER_COLLIDE=0 RR_HEAVY_SOLVER=0 python scripts/time_step.py
echo heavy0-eager1; RR_HEAVY_SOLVER=0 python scripts/time_step.py
