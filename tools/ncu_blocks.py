"""Basic-block view of an ncu `--page source --csv` export: consecutive SASS instructions with
the same execution count are merged; prints address, instructions in block, warp-level
executions, avg active threads, share of all issued instructions, samples."""
import csv, sys, re
rows = list(csv.reader(open(sys.argv[1])))
minshare = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
hi = [i for i, r in enumerate(rows) if r and r[0] == 'Address'][0]
h = rows[hi]
ci, si, smp, ti = h.index('Instructions Executed'), h.index('Source'), h.index('# Samples'), h.index('Thread Instructions Executed')
ins = []
for r in rows[hi + 1:]:
    if r and r[0] == 'Kernel Name':
        break
    try:
        ins.append((r[0], r[si].strip(), int(r[ci]), int(r[ti]), int(r[smp])))
    except Exception:
        pass
tot = sum(i[2] for i in ins)
tots = sum(i[4] for i in ins)
print("total warp instr", tot, "samples", tots)
blocks = []
cur = None
for k, (a, s, n, t, sm) in enumerate(ins):
    if cur is None or n != cur['n'] or re.search(r'\b(BRA|BSYNC|CALL|RET|EXIT)\b', ins[k - 1][1]):
        cur = dict(i0=k, n=n, cnt=0, thr=0, smp=0, first=s)
        blocks.append(cur)
    cur['cnt'] += 1; cur['thr'] += t; cur['smp'] += sm; cur['last'] = s
for b in blocks:
    share = b['n'] * b['cnt'] / tot * 100
    if share >= minshare:
        print(f"@{b['i0']:5d} len {b['cnt']:4d} exec {b['n']:>11d} thr {b['thr'] / max(b['n'] * b['cnt'], 1):5.1f} "
              f"share {share:5.2f}% smp {b['smp'] / tots * 100:5.2f}%  {b['first'][:40]:40s} .. {b['last'][:40]}")
