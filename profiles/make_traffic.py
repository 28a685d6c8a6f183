"""profiles/r1_traffic.json from an `ncu --set full --page raw --csv` export of the streaming kernels: per-launch
DRAM bytes (dram__bytes_read.sum, dram__bytes_write.sum) averaged over the captured launches of each kernel.
Usage: python profiles/make_traffic.py profiles/r1_ncu_full_stream.raw.csv > profiles/r1_traffic.json"""
import collections
import csv
import json
import re
import sys

UNIT = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1.0, 'ms': 1e3}


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    col = {k: hdr.index(k) for k in ('Kernel Name', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum')}
    acc = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0])
    for r in rows[2:]:
        name = r[col['Kernel Name']].split('(')[0].split('::')[-1].replace('void ', '').strip()
        name = re.sub(r'einsum_kernel<(\d)[^>]*>', r'einsum_kernel<\1>', name)   # the STATS variant is not a decoder kernel
        if 'einsum_kernel' in name and re.search(r'einsum_kernel<\d, *1>', r[col['Kernel Name']]):
            name += ' (conv1x1 + statistics)'
        a = acc[name]
        for i, k in enumerate(('dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum')):
            a[i] += float(r[col[k]].replace(',', '')) * UNIT[units[col[k]]]
        a[3] += 1
    out = {n: dict(dram_read_bytes=a[0] / a[3], dram_write_bytes=a[1] / a[3], ncu_duration_us=a[2] / a[3],
                   launches_captured=a[3]) for n, a in acc.items()}
    print(json.dumps(dict(source='%s (ncu --set full --clock-control none, B=4, 128x256 decoder map, per launch)' % path,
                          kernels=out), indent=1))


if __name__ == '__main__':
    main(sys.argv[1])
