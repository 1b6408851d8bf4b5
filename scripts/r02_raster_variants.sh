for v in flat0 flat1 flat2 flat3 flat4 flat6; do RR_LIB=real_robots_b200/variants/lib_$v.so python scripts/time_render.py; done
