"""profiles/dedup_stage_traffic.json from an ncu launch list of one bench.py solve:
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity
    python tools/dedup_traffic.py launches.csv profiles/dedup_stage_traffic.json
DRAM bytes (read + write) per launch of the dedup-stage kernels (thread / warp / CTA kernel of the card-set-grouped
level), averaged over the launches of the LAST solve in the list -- the figure bench.py prints as roofline.traffic."""
import csv
import json
import re
import sys

DEDUP = ('m2_group_tiny_kernel', 'm2_group_warp_kernel', 'm2_group_big_kernel')


def main(src, dst):
    rows = list(csv.reader(open(src, errors='replace')))
    h = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    hdr = rows[h]
    ki, mi, vi, ui, ii = (hdr.index(x) for x in ('Kernel Name', 'Metric Name', 'Metric Value', 'Metric Unit', 'ID'))
    launches = {}
    for r in rows[h + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(',', ''))
        u = r[ui]
        scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(u, 1)
        d = launches.setdefault(int(r[ii]), {'name': re.sub(r'[<(].*', '', r[ki]).split('::')[-1]})
        d[r[mi]] = v * scale
    ids = sorted(launches)
    # the last solve = everything after the last root kernel
    roots = [i for i in ids if launches[i]['name'] == 'm2_root_kernel']
    ids = [i for i in ids if i >= (roots[-1] if roots else 0)]
    per = {}
    for i in ids:
        d = launches[i]
        if d['name'] in DEDUP:
            p = per.setdefault(d['name'], {'launches': 0, 'dram_bytes': 0.0, 'ms': 0.0})
            p['launches'] += 1
            p['dram_bytes'] += d.get('dram__bytes_read.sum', 0.0) + d.get('dram__bytes_write.sum', 0.0)
            p['ms'] += d.get('gpu__time_duration.sum', 0.0)
    n = sum(p['launches'] for p in per.values())
    tot = sum(p['dram_bytes'] for p in per.values())
    out = {'what': 'DRAM bytes (read + write) per launch of the dedup-stage kernels, one beam-30M solve under ncu (cold caches, serialised)',
           'source': src.split('/')[-1], 'launches': n, 'dram_bytes_total': tot, 'dram_bytes_per_launch': tot / max(n, 1),
           'kernel_ms_total_under_ncu': sum(p['ms'] for p in per.values()), 'per_kernel': per}
    json.dump(out, open(dst, 'w'), indent=1)
    print(json.dumps(out))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
