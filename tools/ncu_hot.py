"""Summarise an `ncu --page source --csv` dump: per kernel, the SASS instructions with the most stall samples."""
import csv, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
seen = set()  # the source page lists a kernel once per view: print each (kernel, sample count) once
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name = rows[i][1]
        hdr = rows[i + 1]
        j = i + 2
        body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            if len(rows[j]) == len(hdr):
                body.append(dict(zip(hdr, rows[j])))
            j += 1
        tot = sum(int(r["# Samples"] or 0) for r in body)
        if (name, tot) in seen:
            i = j
            continue
        seen.add((name, tot))
        print(f"== {name[:90]}  total samples {tot}, {len(body)} instrs")
        stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        agg = {c: sum(int(r[c] or 0) for r in body) for c in stall_cols}
        print("   stalls:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot * 0.01})
        for r in sorted(body, key=lambda r: -int(r["# Samples"] or 0))[:top]:
            st = {c[6:]: int(r[c] or 0) for c in stall_cols if int(r[c] or 0) > 0}
            st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
            print(f"   {int(r['# Samples']):6d}  {r['Address'][-5:]}  {r['Source'][:70]:70s} {st}")
        i = j
    else:
        i += 1
