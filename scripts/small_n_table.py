#!/usr/bin/env python3
"""Dev tool: condense the output of scripts/small_n.py (jsonl) into the table kept in profiles/."""
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1])]
for dt in ('float32', 'float64'):
    print(dt)
    for n in sorted({r['n'] for r in rows if r['dtype'] == dt}):
        rr = [r for r in rows if r['dtype'] == dt and r['n'] == n]
        auto = [r for r in rr if r['forced'] == -1][0]
        small = {r['forced']: r['dev_us_per_step'] for r in rr if r['forced'] >= 200}
        old = {r['forced']: r['dev_us_per_step'] for r in rr if 0 <= r['forced'] < 200}
        print(n, 'auto', auto['variant'], auto['dev_us_per_step'], 'wall', auto['wall_us_per_step'], 'G/s', auto['g_inter_s'], 'small', small, 'round-1 best', min(old.values()))
