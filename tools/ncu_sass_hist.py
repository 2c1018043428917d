"""Opcode histogram (dynamic warp instructions + stall samples) from `ncu --page source --csv` output."""
import csv, sys, collections, subprocess, io
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
# may contain several kernels: split on "Kernel Name" rows
k = None
data = collections.OrderedDict()
hdr = None
for r in rows:
    if r and r[0] == "Kernel Name":
        k = r[1][:80]; data[k] = []; hdr = None; continue
    if r and r[0] == "Address":
        hdr = r; continue
    if hdr and len(r) >= len(hdr) - 2:
        data[k].append(r)
for k, rs in data.items():
    iS, iN, iE = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    ops = collections.defaultdict(lambda: [0, 0])
    tot_e = tot_s = 0
    for r in rs:
        src = r[iS].strip()
        parts = src.split()
        op = parts[1] if parts[0].startswith("@") else parts[0]
        op = op.rstrip(";")
        base = ".".join(op.split(".")[:2]) if op.startswith(("LDS", "STS", "LDG", "STG", "HMMA", "LDSM", "ATOM", "RED", "MUFU", "SHFL", "BAR", "LDGSTS")) else op.split(".")[0]
        e, s = int(r[iE] or 0), int(r[iN] or 0)
        ops[base][0] += e; ops[base][1] += s
        tot_e += e; tot_s += s
    print(f"== {k}: {tot_e} warp-instructions, {tot_s} samples")
    for op, (e, s) in sorted(ops.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
        print(f"  {op:14s} {e:12d} {100*e/tot_e:5.1f}% inst   {100*s/max(tot_s,1):5.1f}% samples")
    # hottest individual instructions
    print("  -- hottest instructions by samples")
    for r in sorted(rs, key=lambda r: -int(r[iN] or 0))[:12]:
        print(f"     {100*int(r[iN] or 0)/max(tot_s,1):5.1f}%  {r[iS].strip()[:100]}")
