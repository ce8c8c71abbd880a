for v in 0 1 0 1; do echo "== CSMOE_SCORE_EPILOGUE=$v"; CSMOE_SCORE_EPILOGUE=$v python bench.py --steps 10 | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('router', round(d['ms_per_step'],3), 'competition', round(d['competition']['ms_per_step'],3))"; CSMOE_SCORE_EPILOGUE=$v python scripts/config_sweep.py --steps 10 --only "C5/C2' SigLIP" 2>&1 | tail -1 | cut -d'|' -f4,5; done
