"""Instruction mix / hottest SASS lines of one kernel from an ncu report (run where ncu is installed)."""
import csv, sys, collections, subprocess
rep = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 1; top = int(sys.argv[3]) if len(sys.argv) > 3 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
out = []; hdr = None; k = 0
for r in rows:
    if r and r[0] == 'Kernel Name':
        k += 1
        if k > which: break
        out = []
        continue
    if r and r[0] == 'Address': hdr = r; continue
    if hdr and len(r) == len(hdr): out.append(r)
ie = hdr.index('Instructions Executed'); st = hdr.index('Warp Stall Sampling (All Samples)')
tot = sum(int(r[ie]) for r in out)
nw = max(int(r[ie]) for r in out[:3])
print('total warp instr', tot, 'n sass', len(out), 'warps', nw, 'per warp %.0f' % (tot / nw))
ops = collections.Counter()
for r in out:
    t = r[1].split()
    op = t[0] if not t[0].startswith('@') else t[1]
    ops[op.split('.')[0]] += int(r[ie])
for o, c in ops.most_common(24): print('  %-8s %5.1f%%  %7.0f/warp' % (o, 100 * c / tot, c / nw))
stalls = sum(int(r[st]) for r in out)
print('stall samples', stalls)
if top:
    for r in sorted(out, key=lambda r: -int(r[st]))[:top]: print(r[0][-5:], r[1][:80], r[ie], r[st])
