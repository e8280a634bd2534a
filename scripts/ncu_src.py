"""Top stall sites of an ncu report's source page.  Usage: python scripts/ncu_src.py file.ncu-rep [n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; data = [d for d in rows[2:] if len(d) == len(hdr)]
isrc = hdr.index('Source'); isamp = hdr.index('# Samples'); iex = hdr.index('Instructions Executed')
stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(d[isamp] or 0) for d in data)
print(rows[0][1][:100]); print('total samples', tot)
agg = {}
for d in data:
    for i in stalls:
        agg[hdr[i]] = agg.get(hdr[i], 0) + int(d[i] or 0)
print(sorted(agg.items(), key=lambda kv: -kv[1])[:8])
for d in sorted(data, key=lambda d: -int(d[isamp] or 0))[:n]:
    st = sorted([(int(d[i] or 0), hdr[i][6:]) for i in stalls], reverse=True)[:2]
    print(d[isamp].rjust(7), d[iex].rjust(9), d[isrc][:70].ljust(70), st)
