cd /root/repo
run() { echo -n "$1 | "; shift; env "$@" timeout 120 python tools/gpu_time.py $OBJ $N 40 graph 2>&1 | tail -1 | cut -c1-90; }
for N in 4096 16384; do export N; export OBJ=box
run base HOIC_FUSE=0
run p1pad4k HOIC_FUSE=0 HOIC_P1_PAD=4096
run p3spad3k HOIC_FUSE=0 HOIC_P3S_PAD=3072
run both HOIC_FUSE=0 HOIC_P1_PAD=4096 HOIC_P3S_PAD=3072
run fused HOIC_FUSE=1
done
