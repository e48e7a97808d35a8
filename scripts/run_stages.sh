for st in 4 5 6; do for w in plain qkv up proj down; do echo -n "stages=$st "; SVIT_GEMM_STAGES=$st python scripts/gemm_probe.py $w; done; done
echo "1-CTA kernel:"; for w in plain qkv; do SVIT_GEMM_NO_PAIR=1 python scripts/gemm_probe.py $w; done
