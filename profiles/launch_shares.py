"""One decoder step out of an `ncu --metrics gpu__time_duration.sum --csv` launch list, grouped by kernel.
Usage: python profiles/launch_shares.py profiles/r1_launches_bench.csv [step_index] > profiles/r1_launch_shares.txt"""
import collections
import csv
import re
import sys


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    kn, mv, mu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    out = []
    for r in rows[1:]:
        name = r[kn].split('(')[0].replace('void ', '').split('::')[-1].strip()
        name = re.sub(r'<(\d)[^>]*>', r'<\1>', name)
        ns = float(r[mv].replace(',', '')) * {'ns': 1.0, 'us': 1e3, 'ms': 1e6}[r[mu]]
        out.append((name, ns / 1e3))
    return out


def main(path, step=3):
    ls = launches(path)
    starts = [i for i, (n, _) in enumerate(ls) if n == 'binarise_kernel']
    a, b = starts[step], (starts[step + 1] if step + 1 < len(starts) else len(ls))
    agg = collections.OrderedDict()
    for n, us in ls[a:b]:
        if n.startswith('upsample2x') and 'upsample2x_kernel' in agg and agg['upsample2x_kernel'][0] >= 2:
            break   # the kernel micro-benchmarks that follow the steps
        c = agg.setdefault(n, [0, 0.0])
        c[0] += 1
        c[1] += us
    tot = sum(v[1] for v in agg.values())
    print('one decoder step (#%d of the run) from %s: ncu gpu__time_duration.sum, cold-cache, serialised' % (step + 1, path))
    for n, (c, us) in agg.items():
        print('%-28s x%-4d %8.1f us  %5.1f%%' % (n, c, us, 100 * us / tot))
    print('%-28s x%-4d %8.1f us' % ('total', sum(v[0] for v in agg.values()), tot))


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 3)
