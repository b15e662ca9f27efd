"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: share of time per kernel."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    iK, iV, iM = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for row in r:
        if len(row) <= iV or row[iM] != "gpu__time_duration.sum":
            continue
        name = row[iK]
        v = float(row[iV].replace(",", ""))
        m = re.match(r"(?:void )?(?:b2rl::)?([A-Za-z0-9_]+)(<.*>)?", name)
        short = m.group(1)
        if m.group(2):
            t = m.group(2)
            tile = re.match(r"<(\d+), (\d+), (\d+), (\d+), (\d+)", t)
            ops = re.findall(r"OpTraits<(\d), (\d), (\d), \(bool\)(\d), \(bool\)(\d)>", t)
            epi = re.findall(r"EpiTraits<(\d), (\d), (\d)>", t)
            if tile and ops:
                short += f"<{tile.group(1)}x{tile.group(2)} A{''.join(ops[0])} B{''.join(ops[-1])} E{''.join(epi[0]) if epi else ''}>"
            elif epi:
                short += f"<E{''.join(epi[0])}>"
            else:
                short += t[:30]
        agg[short][0] += 1
        agg[short][1] += v
        n += 1
    tot = sum(v for _, v in agg.values())
    print(f"# {path}: {n} launches, {tot / 1e3:.1f} us total")
    print(f"# {'share':>7} {'count':>6} {'avg_us':>9}  kernel")
    for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{v / tot * 100:7.2f}% {c:6d} {v / c / 1e3:9.1f}  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
