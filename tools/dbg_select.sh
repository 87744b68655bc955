for d in ${DBG_MODES:-0 1 2 3}; do echo "== debug $d"; PCQ_SELECT_DEBUG=$d timeout 200 python tools/sweep.py --points 67108864 --only select 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)
    if d['variant']=='staged' and d['format'] in (1,) : print(d['layout'],d['query'],'ms=%.3f'%d['ms'])
"; done
