"""Top SASS instructions by stall samples from `ncu --page source --csv` output."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = [i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r]
start = hi[-1] if len(hi) > 1 else hi[0]     # last block is SASS view
hdr = rows[start]
body = [r for r in rows[start + 1:] if len(r) == len(hdr)]
si, ni, ii = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed')
ti = hdr.index('Avg. Threads Executed')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[ni] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
ranked = sorted(enumerate(body), key=lambda kv: -int(kv[1][ni] or 0))[:top]
for idx, r in sorted(ranked):
    st = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    print(f"{idx:5d} {int(r[ni] or 0):7d} {100*int(r[ni] or 0)/max(tot,1):5.1f}% ex={r[ii]:>10s} thr={r[ti]:>5s} {r[si][:70]:70s} {st}")
