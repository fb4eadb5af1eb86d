"""Top source lines of an ncu `--page source --print-source cuda,sass --csv` dump by stall samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[2]
col = lambda name: [i for i, h in enumerate(hdr) if h == name][0]
c_samp, c_inst, c_l2g, c_l2l = col('# Samples'), col('Instructions Executed'), col('L2 Theoretical Sectors Global'), col('L2 Theoretical Sectors Local')
stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
out, cur, tot = [], None, 0
for r in rows:
    if r and r[0] == 'File Path':
        cur = r[1].split('/')[-1]; continue
    if len(r) > c_samp and r[0].isdigit() and r[2] == '-':
        try: s = int(r[c_samp])
        except ValueError: continue
        tot += s
        st = sorted([(int(r[i]) if r[i].isdigit() else 0, h[6:]) for i, h in stalls], reverse=True)[:3]
        out.append((s, cur, r[0], r[1].strip()[:70], r[c_inst], r[c_l2g], r[c_l2l], st))
print('total samples', tot)
for o in sorted(out, reverse=True)[:top]: print(o)
