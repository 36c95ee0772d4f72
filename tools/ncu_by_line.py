"""Attribute an ncu SASS source-page CSV of one kernel to CUDA source lines through the line table
of the same cubin (nvdisasm -g -c).  usage: ncu_by_line.py src.csv dis_g.txt kernel_substr [stall] [top]
stall: a stall column ('stall_long_sb', 'stall_no_inst', ...) or '# Samples' (default)."""
import csv
import re
import sys


def line_table(dis_txt, kern):
    tab, sec, cur = {}, None, None
    for l in open(dis_txt):
        m = re.match(r'\s*\.section\s+\.text\.(\S+?),', l)
        if m:
            sec = m.group(1)
            cur = None
            continue
        if not sec or kern not in sec:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        m = re.search(r'/\*([0-9a-f]{4,6})\*/\s+(\S.*?);', l)
        if m:
            tab[int(m.group(1), 16)] = cur
    return tab


def main(src_csv, dis_txt, kern, col='# Samples', top=40):
    tab = line_table(dis_txt, kern)
    data, h, want = [], None, False
    for r in csv.reader(open(src_csv)):
        if r and r[0] == 'Kernel Name':
            want = kern.replace('ILi', '<').split('<')[0].split('EE')[0] in r[1]
            if 'ILi' in kern:
                n = kern.split('ILi')[1].split('E')[0]
                want = want and any(t in r[1] for t in ('<%s>' % n, '<%s,' % n, '<(int)%s>' % n, '<(int)%s,' % n))
            continue
        if r and r[0] == 'Address':
            h = r
            ia, ic, isamp = h.index('Address'), h.index(col), h.index('# Samples')
            continue
        if not want or h is None:
            continue
        try:
            data.append((int(r[ia], 16), int(r[ic] or 0), int(r[isamp] or 0), r[1]))
        except Exception:
            pass
    base = min(d[0] for d in data)
    agg, tot, tots = {}, 0, 0
    for a, v, s, txt in data:
        k = tab.get(a - base)
        d = agg.setdefault(k, [0, 0])
        d[0] += v
        d[1] += s
        tot += v
        tots += s
    print("total %s %d (all samples %d)" % (col, tot, tots))
    src = {}
    for k, d in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(top)]:
        txt = ''
        if k:
            if k[0] not in src:
                try:
                    src[k[0]] = open('/root/repo/gymcity_b200/csrc/' + k[0]).read().split('\n')
                except Exception:
                    src[k[0]] = []
            if 0 < k[1] <= len(src[k[0]]):
                txt = src[k[0]][k[1] - 1].strip()[:90]
        print("%5.1f%% (%5.1f%% of samples) %s:%s  %s" % (100.0 * d[0] / max(1, tot), 100.0 * d[1] / max(1, tots),
                                                   k[0] if k else '?', k[1] if k else '?', txt))


if __name__ == "__main__":
    main(*sys.argv[1:])
