"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: share of time per kernel."""
import collections
import csv
import re
import sys


# (A traits, B traits, epilogue traits) -> what the launch is  [elem,rowmap,redmap,red_fast,ones]
ROLES = {
    (("1", "1", "2", "1", "0"), ("0", "0", "0", "1", "0"), ("0", "1", "0")): "conv fwd, uint8 frames (layer 1)",
    (("0", "1", "2", "1", "0"), ("0", "0", "0", "1", "0"), ("0", "1", "0")): "conv fwd, fp32 activations",
    (("0", "0", "0", "1", "0"), ("0", "0", "0", "1", "0"), ("0", "0", "0")): "linear fwd",
    (("0", "0", "0", "1", "0"), ("0", "0", "0", "0", "0"), ("0", "0", "0")): "linear dX",
    (("0", "0", "0", "0", "0"), ("0", "0", "0", "0", "1"), ("2", "0", "0")): "linear dW+db",
    (("1", "2", "1", "1", "1"), ("0", "0", "1", "1", "0"), ("3", "0", "0")): "conv wgrad+bgrad, uint8 frames (layer 1)",
    (("0", "2", "1", "1", "1"), ("0", "0", "1", "1", "0"), ("3", "0", "0")): "conv wgrad+bgrad, fp32 activations",
    (("0", "1", "0", "0", "0"), ("0", "0", "0", "0", "0"), ("1", "1", "2")): "conv dgrad (col2im atomics)",
    (("0", "0", "1", "1", "0"), ("1", "2", "1", "0", "1"), ("2", "0", "0")): "conv wgrad (old roles), uint8",
    (("0", "0", "1", "1", "0"), ("0", "2", "1", "0", "1"), ("2", "0", "0")): "conv wgrad (old roles), fp32",
}


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
        m = re.match(r"(?:void )?(?:b2rl::)?([A-Za-z0-9_]+)", name)
        short = m.group(1)
        tile = re.search(r"igemm_kernel<(\d+), (\d+), (\d+), (\d+), (\d+)", name)
        ops = re.findall(r"OpTraits<(\d), (\d), (\d), (\d), (\d)>", name)
        epi = re.findall(r"EpiTraits<(\d), (\d), (\d)>", name)
        if tile and ops:
            role = ROLES.get((ops[0], ops[1], epi[0] if epi else None), "")
            short += f"<{tile.group(1)}x{tile.group(2)} A={''.join(ops[0])} B={''.join(ops[1])} E={''.join(epi[0])}> {role}"
        elif epi:
            short += f"<E={''.join(epi[0])}>"
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
