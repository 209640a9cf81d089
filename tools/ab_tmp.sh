export TUTU_LIB=/root/repo/tuturenderer_b200/libtutu_b200_exp.so
run() { env "$@" timeout 120 python tools/gpu_bdpt_scenes.py 2>&1 | tail -1; }
run TUTU_QUEUE_LANES=0
run TUTU_QUEUE_LANES=3 TUTU_LEAF_BATCH=16
run TUTU_QUEUE_LANES=2 TUTU_LEAF_BATCH=16
run TUTU_QUEUE_LANES=3 TUTU_LEAF_BATCH=8
run TUTU_QUEUE_LANES=3 TUTU_LEAF_BATCH=20
run TUTU_QUEUE_LANES=0
