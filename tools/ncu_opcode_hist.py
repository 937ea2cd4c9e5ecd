"""Executed warp-instructions per opcode and per warp butterfly from an ncu report's SASS source page:
    ncu -i rep.ncu-rep --page source --csv --print-source sass > src.csv
    python tools/ncu_opcode_hist.py src.csv <polys> <logN> <stages of kernel 1> <stages of kernel 2> ...
(the opcode mix the dispatch model of DESIGN.md section 5b is fed with)."""
import collections, csv, re, sys

rows = list(csv.reader(open(sys.argv[1])))
polys, logN = int(sys.argv[2]), int(sys.argv[3])
stages = [int(a) for a in sys.argv[4:]]
kern, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        kern.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and r:
        cur["rows"].append(r)
seen = set()
ki = 0
for k in kern:
    if k["name"] in seen:
        continue
    seen.add(k["name"])
    h = k["hdr"]
    iS, iE, iSm = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    cnt, smp, tot = collections.Counter(), collections.Counter(), 0
    for r in k["rows"]:
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iS].strip())
        op = re.sub(r"^([A-Z0-9]+(\.WIDE)?(\.HI)?(\.X)?).*", r"\1", m.group(2) if m else r[iS].strip())
        e = int(r[iE]); cnt[op] += e; tot += e; smp[op] += int(r[iSm])
    nb = polys * (1 << (logN - 1)) * stages[ki] / 32
    ki += 1
    print(f"{k['name'][:110]}\n  {tot} warp-instructions, {tot / nb:.1f} per warp butterfly ({stages[ki - 1]} stages)")
    ts = sum(smp.values())
    for op, c in cnt.most_common(24):
        print(f"    {op:14s} {c / nb:7.2f} per warp butterfly   stall samples {100 * smp[op] / ts:5.1f} %")
