"""profiles/*.md from an ncu launch list (ncu --metrics gpu__time_duration.sum --csv): per-kernel totals
and shares of ONE step of bench.py, delimited by consecutive roi_align_bwd launches."""
import collections
import csv
import sys


def main(path, which=2, title="", command=""):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if r and r[0] == 'ID':
            hdr, start = r, i + 1
            break
    i_n, i_v, i_m = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
    launches = [(r[i_n], float(r[i_v].replace(',', ''))) for r in rows[start:]
                if len(r) > i_v and r[i_m] == 'gpu__time_duration.sum']
    idx = [i for i, (n, v) in enumerate(launches) if 'roi_align_bwd' in n]
    which = int(which)
    a, b = idx[which] + 1, idx[which + 1] + 1
    step = launches[a:b]
    tot = sum(v for n, v in step)
    agg = collections.OrderedDict()
    for n, v in step:
        k = n.replace('void ', '').replace('coin::', '')[:72]
        agg.setdefault(k, [0.0, 0])
        agg[k][0] += v
        agg[k][1] += 1
    print(f"# {title}\n")
    print(f"Command: `{command}`")
    print("(per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes).\n")
    print(f"Launches in the step: {len(step)}; summed kernel time {tot / 1e3:.1f} us.\n")
    print("| us | share | launches | kernel |\n|---:|---:|---:|---|")
    for k, (v, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
        print(f"| {v / 1e3:.1f} | {100 * v / tot:.1f}% | {c} | `{k}` |")


if __name__ == "__main__":
    main(*sys.argv[1:])
