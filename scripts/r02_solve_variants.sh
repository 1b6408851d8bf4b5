for v in smb6 s5w4 s4w4 s12w2 s24w1 s10w2 s20w1; do RR_LIB=real_robots_b200/variants/lib_$v.so python scripts/time_step.py; done
