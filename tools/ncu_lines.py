"""Aggregate an `ncu -i X.ncu-rep --page source --csv --print-source cuda[,sass]` export by CUDA source line,
per profiled launch (the export lists the launches one after another, each as a run of per-file sections).

    python tools/ncu_lines.py src.csv [top=40] [launch index, default all]
"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
only = int(sys.argv[3]) if len(sys.argv) > 3 else None

launches = []  # (function name, [(inst, samples, thread inst, file, line, text, long_sb)])
fn, hdr, cur_fn, seen_files = None, None, None, set()
for r in rows:
    if len(r) == 2 and r[0] in ("File Path", "File Name"):
        fn = r[1].split('/')[-1]
        continue
    if len(r) == 2 and r[0] == "Function Name":
        if cur_fn != r[1] or fn in seen_files:
            launches.append((r[1], []))
            cur_fn, seen_files = r[1], set()
        seen_files.add(fn)
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) >= 10 and r[0] not in ("", "Line No"):
        try:
            inst = int(r[hdr.index("Instructions Executed")])
            smp = int(r[hdr.index("# Samples")])
            tinst = int(r[hdr.index("Thread Instructions Executed")])
        except ValueError:
            continue
        if not launches:
            launches.append(("?", []))
        launches[-1][1].append((inst, smp, tinst, fn, int(r[0]), r[1].strip()[:110]))

for idx, (name, agg) in enumerate(launches):
    if only is not None and idx != only:
        continue
    tot = sum(a[0] for a in agg) or 1
    tots = sum(a[1] for a in agg) or 1
    print("==== launch %d: %s" % (idx, name[:90]))
    print("total warp inst %d  samples %d  thr/inst %.1f" % (tot, tots, sum(a[2] for a in agg) / tot))
    byfile = collections.defaultdict(lambda: [0, 0, 0])
    for a in agg:
        byfile[a[3]][0] += a[0]
        byfile[a[3]][1] += a[1]
        byfile[a[3]][2] += a[2]
    for k, v in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
        print("%-28s inst %5.1f%%  samples %5.1f%%  thr/inst %.1f" % (k, 100 * v[0] / tot, 100 * v[1] / tots, v[2] / max(v[0], 1)))
    print()
    for a in sorted(agg, key=lambda a: -a[0])[:top]:
        print("%5.2f%% inst %5.2f%% smp  %4.1f thr  %s:%d  %s" % (100 * a[0] / tot, 100 * a[1] / tots, a[2] / max(a[0], 1), a[3], a[4], a[5]))
    print()
