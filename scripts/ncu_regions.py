"""Stall samples of the brick kernel by code region, from `ncu --page source --csv` output (developer helper)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r]
hdr = rows[hi[-1]]; body = [r for r in rows[hi[-1] + 1:] if len(r) == len(hdr)]
ni, si, ii = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
S = [int(r[ni] or 0) for r in body]; EX = [int(r[ii] or 0) for r in body]
tot = sum(S)
cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
# the walk = the instructions executed as often as the first MUFU.RCP64H
mu = [i for i, r in enumerate(body) if 'MUFU.RCP64H' in r[si]]
walk_ex = EX[mu[0]]
walk = [i for i in range(len(body)) if EX[i] == walk_ex]
w0, w1 = walk[0], walk[-1]
spin = [i for i, r in enumerate(body) if 'NANOSLEEP' in r[si]]
def show(name, a, b):
    s = sum(S[a:b]); agg = {h: sum(int(r[i] or 0) for r in body[a:b]) for i, h in cols}
    t = max(sum(agg.values()), 1)
    print(f"{name:34s} {s:6d} {100*s/tot:5.1f}%  inst {sum(EX[a:b]):>11d}  ", {k: round(100*v/t) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:5]})
print("total samples", tot, "walk range", w0, w1, "spins at", spin)
marks = [0] + [x for sp in spin for x in (sp - 1, sp + 6)] + [w0, w1 + 1, len(body)]
marks = sorted(set(m for m in marks if 0 <= m <= len(body)))
for a, b in zip(marks, marks[1:]):
    show(f"[{a},{b}) {body[a][si][:22]}", a, b)
