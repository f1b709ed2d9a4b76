"""Per-instruction stall breakdown of an .ncu-rep region: python tools/ncu_stalls.py rep [min_pct]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
minp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hh, data = rows[1], rows[2:]
isrc, ist = hh.index('Source'), hh.index('Warp Stall Sampling (All Samples)')
reasons = [h for h in hh if h.startswith('stall_') and 'Not Issued' not in h]
idx = {r: hh.index(r) for r in reasons}
tot = sum(int(x[ist]) for x in data) or 1
print('total samples', tot)
for i, x in enumerate(data):
    n = int(x[ist])
    if n >= tot * minp / 100:
        rs = sorted(((int(x[idx[r]] or 0), r) for r in reasons), reverse=True)[:3]
        print('%5d %-70s %5.1f%%  %s' % (i, x[isrc].strip()[:70], 100 * n / tot, ' '.join('%s=%d' % (r[6:], c) for c, r in rs if c)))
