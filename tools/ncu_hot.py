"""Groups the SASS of an ncu source page (ncu -i X --page source --csv > file) into blocks of equal
execution count and prints each block's share of executed instructions and stall samples."""
import csv
import sys


def main(path, thresh=0.01):
    rows = list(csv.reader(open(path)))
    hdr, data = rows[1], rows[2:]
    i_s, i_e, i_w = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
    tot = sum(int(r[i_e]) for r in data)
    totw = sum(int(r[i_w]) for r in data)
    print("instructions executed", tot, "samples", totw)
    blk, cur = [], None
    for i, r in enumerate(data):
        e, w = int(r[i_e]), int(r[i_w])
        if cur and abs(e - cur['e']) <= 0.02 * max(e, cur['e'], 1):
            cur['n'] += 1; cur['sum'] += e; cur['w'] += w; cur['end'] = i
        else:
            cur = {'start': i, 'end': i, 'e': e, 'n': 1, 'sum': e, 'w': w}
            blk.append(cur)
    for b in blk:
        if b['sum'] > float(thresh) * tot:
            ops = {}
            for r in data[b['start']:b['end'] + 1]:
                t = r[i_s].split()
                op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]
                ops[op] = ops.get(op, 0) + 1
            print(f"[{b['start']:4d}-{b['end']:4d}] n={b['n']:3d} exec/instr={b['e']:9d} share={b['sum']/tot*100:5.1f}% "
                  f"stall={b['w']/totw*100:5.1f}%", dict(sorted(ops.items(), key=lambda x: -x[1])[:9]))


if __name__ == "__main__":
    main(*sys.argv[1:])
