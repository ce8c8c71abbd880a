for d in 0 1 2 3; do echo "== DBG=$d"; CSMOE_GEMM_DBG=$d CSMOE_GEMM_WIDE=0 CSMOE_GEMM_EPI=direct python scripts/gemm_bench.py 20 2>&1 | grep -E "fwd1 plain|fwd2|dgrad1|wgrad1"; done
