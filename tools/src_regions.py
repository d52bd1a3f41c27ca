"""Per-instruction view of an ncu report (--import-source on): warp-level executions, active lanes and stall samples.
usage: python tools/src_regions.py REPORT.ncu-rep [min_share_pct]   (prints the hottest SASS lines + totals)"""
import csv, subprocess, sys, io
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.6
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
blocks = []; cur = None
for r in rows:
    if r and r[0] == 'Kernel Name': cur = []; blocks.append((r[1], cur)); continue
    if cur is not None and r: cur.append(r)
name, b = blocks[0]; h = b[0]; data = b[1:]; ix = {n: i for i, n in enumerate(h)}
I = lambda r, k: int(r[ix[k]] or 0)
ti = sum(I(r, 'Instructions Executed') for r in data); tt = sum(I(r, 'Thread Instructions Executed') for r in data); ts = sum(I(r, '# Samples') for r in data)
print(name[:90]); print("SASS lines %d  warp-inst %.1fM  lanes/inst %.2f  samples %d" % (len(data), ti / 1e6, tt / ti, ts))
base = int(data[0][0], 16)
for r in data:
    ie = I(r, 'Instructions Executed'); s = I(r, '# Samples')
    if 100.0 * s / ts >= thr or '--all' in sys.argv:
        print("%04x %-50s exec %6.2fM lanes %5.1f smp %4.1f%% lsb %s" % (int(r[0], 16) - base, r[1].strip()[:50], ie / 1e6,
              I(r, 'Thread Instructions Executed') / max(ie, 1), 100.0 * s / ts, r[ix['stall_long_sb']]))
