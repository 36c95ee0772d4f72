"""Dynamic call counts of one kernel from an ncu SASS source-page CSV: which out-of-line device functions are entered
how often (warp-level executions of the CALL instructions, by callee), with the callee names from the symbol table of
the same library.  usage: ncu_calls.py src.csv lib.so kernel_substr [top]"""
import bisect
import collections
import csv
import re
import subprocess
import sys


def symbols(lib, kern):
    elf = subprocess.run(["cuobjdump", "-elf", lib], capture_output=True, text=True).stdout
    rows = []
    for l in elf.splitlines():
        m = re.match(r'\s*0x[0-9a-f]+\s+(0x[0-9a-f]+)\s+(0x[0-9a-f]+)\s+(0x[0-9a-f]+)\s+\S+\s+(0x[0-9a-f]+)\s+(\S+)\s*$', l)
        if m:
            rows.append((int(m.group(1), 16), int(m.group(2), 16), m.group(4), m.group(5)))
    sec = next(sc for a, sz, sc, name in rows if kern in name and name.startswith('$'))
    syms = sorted((a, sz, name.split('$')[-1]) for a, sz, sc, name in rows if sc == sec and name.startswith('$'))
    return syms


def main(src_csv, lib, kern, top=40):
    syms = symbols(lib, kern)
    starts = [s[0] for s in syms]

    def fn(off):
        i = bisect.bisect_right(starts, off) - 1
        return '<kernel>' if i < 0 or off >= syms[i][0] + syms[i][1] else syms[i][2]

    want, base, rows = False, None, []
    for r in csv.reader(open(src_csv)):
        if r and r[0] == 'Kernel Name':
            want = kern.split('ILi')[0] in r[1] and (('<(int)%s,' % kern.split('ILi')[1].split('E')[0]) in r[1] or ('<%s,' % kern.split('ILi')[1].split('E')[0]) in r[1]) if 'ILi' in kern else kern in r[1]
            base = None
            continue
        if not want or not r or r[0] == 'Address':
            continue
        try:
            a, ex = int(r[0], 16), int(r[5])
        except ValueError:
            continue
        if base is None:
            base = a
        rows.append((a - base, r[1].strip(), ex))
        if len(rows) > 10 ** 6:
            break
    total = sum(ex for _, _, ex in rows)
    by_callee, by_edge = collections.Counter(), collections.Counter()
    for off, txt, ex in rows:
        m = re.search(r'CALL\S*\s+(0x[0-9a-f]+)', txt)
        if m and base is not None:
            t = int(m.group(1), 16) - base
            by_callee[fn(t)] += ex
            by_edge[(fn(off), fn(t))] += ex
    calls = sum(by_callee.values())
    print("warp-instructions %d, calls %d (one per %.0f instructions)" % (total, calls, total / max(calls, 1)))
    for name, c in by_callee.most_common(int(top)):
        print("%6.2f%% of calls  %9d  %s" % (100.0 * c / calls, c, name))
    print("---- caller -> callee")
    for (a, b), c in by_edge.most_common(int(top)):
        print("%6.2f%%  %9d  %s -> %s" % (100.0 * c / calls, c, a[:40], b[:40]))


if __name__ == "__main__":
    main(*sys.argv[1:])
