"""Summarise an `ncu --page source --csv` dump (SASS view): stall reasons overall, instruction mix, hottest instructions."""
import collections
import csv
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    tot = collections.Counter()
    mix = collections.Counter()
    insts = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        samples = int(r[idx["# Samples"]] or 0)
        ex = int(r[idx["Instructions Executed"]] or 0)
        op = r[idx["Source"]].strip().split()
        op = [o for o in op if not o.startswith("@")]
        name = op[0].split(".")[0] if op else "?"
        mix[name] += ex
        for s in stall_cols:
            tot[s] += int(r[idx[s]] or 0)
        insts.append((samples, ex, r[idx["Source"]].strip(), {s: int(r[idx[s]] or 0) for s in stall_cols if int(r[idx[s]] or 0)}))
    ts = sum(tot.values())
    print("# stall samples (all): %d" % ts)
    for s, v in tot.most_common(10):
        print("  %-24s %6.1f%%" % (s, 100.0 * v / ts))
    te = sum(mix.values())
    print("# instruction mix (warp-level executed): %d" % te)
    for s, v in mix.most_common(18):
        print("  %-12s %6.1f%%  %d" % (s, 100.0 * v / te, v))
    print("# hottest instructions")
    for smp, ex, src, st in sorted(insts, key=lambda x: -x[0])[:top]:
        print("  %6d %9d  %-60s %s" % (smp, ex, src[:60], dict(sorted(st.items(), key=lambda x: -x[1])[:3])))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
