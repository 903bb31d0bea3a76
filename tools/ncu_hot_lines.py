"""Hot source lines of one kernel from an ncu report:
    ncu -i rep.ncu-rep --page source --csv --kernel-name regex:<kernel> > src.csv
    python tools/ncu_hot_lines.py src.csv [threshold=0.004]
Prints the stall-reason totals and every source line with more than `threshold` of the samples (cumulative share, samples,
instructions executed, dominant stall reasons)."""
import csv
import sys


def main():
    f = sys.argv[1]
    thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
    rows = list(csv.reader(open(f)))
    hdr = rows[1]
    si, so, ie = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
    data = [r for r in rows[2:] if len(r) == len(hdr) and r[si].isdigit()]
    tot = sum(int(r[si]) for r in data)
    print('rows', len(data), 'total samples', tot, 'total inst', sum(int(r[ie]) for r in data))
    names = [n for n in hdr if n.startswith('stall_') and not n.endswith(')')]
    col = {n: hdr.index(n) for n in names}
    print({n: sum(int(r[col[n]]) for r in data) for n in names})
    cum = 0
    for idx, r in enumerate(data):
        s = int(r[si])
        cum += s
        if s > tot * thr:
            st = ' '.join("%s=%s" % (n[6:], r[col[n]]) for n in names if int(r[col[n]]) > s * 0.15)
            print("%6d cum%5.1f%% %-60s samp %6d exec %8s %s" % (idx, 100 * cum / tot, r[so][:60].strip(), s, r[ie], st))


if __name__ == "__main__":
    main()
