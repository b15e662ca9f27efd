"""Group the SASS of one kernel in an .ncu-rep into runs of equal execution count and print the heavy
runs (instructions executed, stall samples, opcode mix).  usage: ncu_regions.py report.ncu-rep <kernel substring> [min]"""
import csv, subprocess, sys, io, collections

def main(path, sub, thresh=60000):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    kern, data, hdr = None, collections.OrderedDict(), None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            kern = r[1] + f"#{len(data)}"; data[kern] = []; continue
        if r and r[0] == 'Address':
            hdr = r; continue
        if kern: data[kern].append(r)
    for k, v in data.items():
        if sub not in k: continue
        ie, si = hdr.index('Instructions Executed'), hdr.index('# Samples')
        tot = sum(int(r[ie]) for r in v); ts = sum(int(r[si]) for r in v)
        print(k[:100], len(v), 'sass;', tot, 'warp-instr;', ts, 'samples')
        runs = []
        for i, r in enumerate(v):
            c = int(r[ie]); op = r[1].split()[1] if r[1].strip().startswith('@') else r[1].split()[0]
            if runs and runs[-1][1] == c: runs[-1][2] += 1; runs[-1][3] += int(r[si]); runs[-1][4].append(op)
            else: runs.append([i, c, 1, int(r[si]), [op]])
        for st, c, n, s, ops in runs:
            if c * n > thresh or s > ts * 0.02:
                cnt = collections.Counter(o.split('.')[0] for o in ops)
                print(f"  @{st:5d} x{c:7d} n={n:4d} instr={c*n:9d} ({100*c*n/tot:4.1f}%) samples={s:5d} ({100*s/max(ts,1):4.1f}%) {dict(cnt.most_common(7))}")
        break

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 60000)
