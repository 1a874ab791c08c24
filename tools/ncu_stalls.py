"""Summarise an `ncu --page source --csv` export into a markdown stall table.

usage: python tools/ncu_stalls.py <source.csv> [title] > profiles/<name>_stalls.md

Works on the SASS view (default) and on `--print-source cuda` exports (per source line).  The
CSV holds one row per instruction / line with `# Samples` and one column per stall reason.
"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    rows = list(csv.reader(open(path)))
    # the header is the first row that has a '# Samples' column
    hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
    hdr = rows[hi]
    si = hdr.index("Source")
    ni = hdr.index("# Samples")
    ei = hdr.index("Instructions Executed") if "Instructions Executed" in hdr else None
    stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data, cat, tot = [], collections.Counter(), 0
    for r in rows[hi + 1:]:
        try:
            n = int(r[ni])
        except (ValueError, IndexError):
            continue
        tot += n
        st = {hdr[i][6:]: int(r[i]) for i in stall if r[i] not in ("", "0")}
        for k, v in st.items():
            cat[k] += v
        data.append((n, r, st))
    data.sort(key=lambda x: -x[0])
    out = [f"# {title} - warp-stall sampling", "", f"Total samples {tot}.", "",
           "| reason | samples | share |", "|---|---|---|"]
    allst = sum(cat.values()) or 1
    for k, v in cat.most_common():
        out.append(f"| {k} | {v} | {v / allst:.1%} |")
    out += ["", "Top rows by samples:", "", "| samples | executed | source | top stall reasons |",
            "|---|---|---|---|"]
    for n, r, st in data[:40]:
        top = ", ".join(f"{k} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:3])
        ex = r[ei] if ei is not None else ""
        out.append(f"| {n} | {ex} | `{r[si].strip()[:90]}` | {top} |")
    print("\n".join(out))


if __name__ == "__main__":
    main()
