"""Attribute an ncu source-page CSV (SASS view) of one kernel to the device functions inside it,
using nvdisasm's labels of the same cubin.  usage: ncu_by_function.py src.csv dis.txt kernel_substr [section_substr]"""
import csv
import re
import sys


def main(src_csv, dis_txt, kern, sect=None):
    sect = sect or kern          # mangled-name substring of the cubin section (template instantiations)
    labels, sec, off = [], None, 0
    for l in open(dis_txt):
        m = re.match(r'\s*\.section\s+\.text\.(\S+?),', l)
        if m:
            sec = m.group(1)
            continue
        if not sec or sect not in sec:
            continue
        m = re.match(r'^(\$?[_a-zA-Z][\w\$]*):\s*$', l)
        if m and not m.group(1).startswith('.L'):
            labels.append([m.group(1).split('$')[-1], None])
            continue
        m = re.search(r'/\*([0-9a-f]{4,6})\*/', l)
        if m and labels and labels[-1][1] is None:
            labels[-1][1] = int(m.group(1), 16)
    labels = [l for l in labels if l[1] is not None]
    # the CSV holds one table per profiled launch ("Kernel Name" row, header row, SASS rows): launches
    # of the wanted kernel are summed
    data, h, want = [], None, False
    for r in csv.reader(open(src_csv)):
        if r and r[0] == 'Kernel Name':
            want = kern in r[1]
            continue
        if r and r[0] == 'Address':
            h = r
            ia, isamp, iex = h.index('Address'), h.index('# Samples'), h.index('Instructions Executed')
            stall_cols = [i for i, n in enumerate(h) if n.startswith('stall_') and 'Not Issued' not in n]
            continue
        if not want or h is None:
            continue
        try:
            data.append((int(r[ia], 16), int(r[isamp]), int(r[iex]), {h[i]: r[i] for i in stall_cols}))
        except Exception:
            pass
    base = min(d[0] for d in data)
    agg = {}
    for a, s, x, r in data:
        o = a - base
        name = None
        for n, st in labels:
            if st <= o:
                name = n
            else:
                break
        d = agg.setdefault(name, [0, 0, {}])
        d[0] += s
        d[1] += x
        for k, v in r.items():
            try:
                v = int(v)
            except Exception:
                v = 0
            if v:
                d[2][k] = d[2].get(k, 0) + v
    ts, tx = sum(d[0] for d in agg.values()), sum(d[1] for d in agg.values())
    print("total samples %d  warp-instructions %d" % (ts, tx))
    for n, d in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
        top = sorted(d[2].items(), key=lambda kv: -kv[1])[:3]
        print("%5.1f%% samples %5.1f%% inst  %-34s %s" % (100.0 * d[0] / ts, 100.0 * d[1] / tx, str(n)[:34],
                                                       " ".join("%s=%d%%" % (k[6:], 100 * v / max(1, d[0])) for k, v in top)))


if __name__ == "__main__":
    main(*sys.argv[1:5])
