python tools/poisson_sweep.py topo 256 2>/dev/null | python -c "
import json,sys
print(' '.join('%s %.3f' % (r['topology'], r['ms_per_solve']) for r in json.load(sys.stdin)))"
