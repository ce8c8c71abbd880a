for v in 0 1 0 1; do echo "== CSMOE_ROUTER_GEMM=$v"; CSMOE_ROUTER_GEMM=$v python scripts/config_sweep.py --steps 20 --only "C4" 2>&1 | tail -2 | cut -d"|" -f4,5; done
CSMOE_ROUTER_GEMM=1 python scripts/config_sweep.py --steps 5 --only "C4" --profile 2>&1 | grep -v Warn | grep -A24 "router step" | cut -c1-150
