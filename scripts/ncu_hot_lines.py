"""Top stall locations of one kernel from an `ncu --set full --import-source on` report:
    ncu -i prof.ncu-rep --page source --csv --kernel-name regex:NAME > src.csv ; python scripts/ncu_hot_lines.py src.csv [N [launch]]
Prints the N SASS instructions with the most warp-stall samples (index, samples, dominant stall reason, text)."""
import csv
import sys


def main(path, n=25):
    rows = list(csv.reader(open(path)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0          # several launches match: pick one
    rows = rows[starts[which]:starts[which + 1]]
    hdr, data = rows[1], rows[2:]
    isrc, isamp = hdr.index("Source"), hdr.index("# Samples")
    stalls = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[isamp]) for r in data if len(r) > isamp)
    print(f"{rows[0][1][:90]}: {tot} samples, {len(data)} instructions")
    top = sorted(range(len(data)), key=lambda i: -int(data[i][isamp]))[:n]
    for i in sorted(top):
        r = data[i]
        why = max(stalls, key=lambda s: int(r[s[0]] or 0))
        print(f"{i:5d} {int(r[isamp]):6d} {100 * int(r[isamp]) / max(tot, 1):5.1f}%  {why[1][6:]:14s} {r[isrc].strip()[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
