"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list
(second half of the launches = the timed repetition).  usage: python tools/launchsum.py <launches.csv> [all]"""
import csv, collections, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, mi, vi, ii = hdr.index('Kernel Name'), hdr.index('Metric Name'), hdr.index('Metric Value'), hdr.index('ID')
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault(r[ii], {'k': r[ki][:70]})[r[mi]] = float(r[vi].replace(',', ''))
ids = list(d)
if len(sys.argv) < 3:
    ids = ids[len(ids) // 2:]
agg = collections.OrderedDict()
for i in ids:
    x = d[i]
    a = agg.setdefault(x['k'], [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += x.get('gpu__time_duration.sum', 0); a[2] += x.get('dram__bytes_read.sum', 0); a[3] += x.get('dram__bytes_write.sum', 0)
tot = sum(a[1] for a in agg.values())
for k, a in agg.items():
    print('%-70s n=%3d  avg %7.1f us  share %4.1f%%  rd %5.0f MB  wr %5.0f MB' % (k, a[0], a[1] / a[0] / 1e3, 100 * a[1] / tot, a[2] / a[0] / 1e6, a[3] / a[0] / 1e6))
