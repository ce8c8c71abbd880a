for e in staged direct; do echo "== CSMOE_GEMM_EPI=$e"; CSMOE_GEMM_EPI=$e python scripts/gemm_bench.py 10 siglip 2>&1 | grep "fc1"; done
