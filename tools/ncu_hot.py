"""Hot spots from `ncu --page source --csv`: top SASS lines by stall samples + totals per stall reason."""
import csv
import sys


def main(path, top=25):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    data = []
    for r in rows[2:]:  # first kernel section only
        if r and r[0] == "Kernel Name":
            break
        if len(r) == len(hdr) and r[0] != "Address":
            data.append(r)
    iS, iSrc, iN, iInst = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Address"), hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(d[iS] or 0) for d in data)
    print("total samples", tot, "SASS lines", len(data), "warp-instr", sum(int(d[iInst] or 0) for d in data))
    agg = {hdr[i]: sum(int(d[i] or 0) for d in data) for i in stall_cols}
    print("stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
    order = sorted(range(len(data)), key=lambda k: -int(data[k][iS] or 0))[:top]
    for k in sorted(order):
        d = data[k]
        st = {hdr[i][6:]: int(d[i] or 0) for i in stall_cols if int(d[i] or 0)}
        print(f"{k:5d} {int(d[iS]):6d} {100*int(d[iS])/max(tot,1):5.1f}% x{d[iInst]:>8s} {d[iSrc].strip()[:70]:70s} {st}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
