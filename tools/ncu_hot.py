"""Top SASS instructions by warp-stall samples for one kernel of an ncu report (needs --import-source on).
   python tools/ncu_hot.py gpurun_out/prof.ncu-rep "gemm_i8_tcgen05_kernel<(int)256, (int)3, (int)2" [N]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for line in out.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = {"name": line, "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(line)
for b in blocks:
    if pat not in b["name"]:
        continue
    rows = list(csv.reader(io.StringIO("\n".join(b["rows"]))))
    hdr = rows[0]
    iS, iSrc, iEx = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = [r for r in rows[1:] if len(r) > iS and r[iS].isdigit()]
    tot = sum(int(r[iS]) for r in data)
    print(b["name"][:110], "total samples", tot, "instructions", len(data))
    for idx, r in sorted(enumerate(data), key=lambda kv: -int(kv[1][iS]))[:N]:
        st = sorted(((hdr[i][6:], int(r[i])) for i in stall_cols if r[i].isdigit() and int(r[i]) > 0), key=lambda kv: -kv[1])[:3]
        print("%5d  %5.1f%%  exec=%-8s %-60s %s" % (idx, 100.0 * int(r[iS]) / max(tot, 1), r[iEx], r[iSrc].strip()[:60], st))
    break
