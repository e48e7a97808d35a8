for f in 0 1; do for w in proj down; do echo -n "flags=$f "; SVIT_GEMM_FLAGS=$f python scripts/gemm_probe.py $w; done; done
