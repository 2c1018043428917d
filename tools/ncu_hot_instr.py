"""Hottest SASS instructions (stall samples + top stall reasons) of ONE kernel of an .ncu-rep captured with
--import-source on:  python tools/ncu_hot_instr.py report.ncu-rep <kernel index> [top N]"""
import csv, io, subprocess, sys
rep, which = sys.argv[1], int(sys.argv[2])
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
sections, cur = [], None
for r in csv.reader(io.StringIO(txt)):
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "rows": []}
        sections.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and len(r) >= len(cur["hdr"]) - 2:
        cur["rows"].append(r)
print(len(sections), "kernels:", [s["name"][40:90] for s in sections])
s = sections[which]
hdr, data = s["hdr"], s["rows"]
iS, iN, iE = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iN] or 0) for r in data)
print(s["name"][:120]); print("samples", tot, "warp instructions", sum(int(r[iE] or 0) for r in data))
agg = {hdr[i]: sum(int(r[i] or 0) for r in data) for i in stalls}
print("stall reasons:", [(k, f"{100 * v / tot:.1f}%") for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]])
top = sorted(range(len(data)), key=lambda i: -int(data[i][iN] or 0))[:topn]
for i in sorted(top):
    r = data[i]
    st = sorted(((hdr[j], int(r[j] or 0)) for j in stalls if int(r[j] or 0) > 0), key=lambda kv: -kv[1])[:3]
    print(f"{i:5d} {100 * int(r[iN]) / tot:5.2f}%  x{r[iE]:>8s}  {r[iS].strip()[:64]:64s} {st}")
