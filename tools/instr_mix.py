"""Dynamic SASS instruction mix per kernel of an ncu report captured with --import-source on (executed warp instructions
and warp-stall samples per opcode):   python tools/instr_mix.py gpurun_out/hot.ncu-rep [top N] > profiles/instr_mix.md"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
N = int(sys.argv[2]) if len(sys.argv) > 2 else 14
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for line in out.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = {"name": line, "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(line)
print("# Dynamic instruction mix (`ncu --set full --import-source on`, source page)\n\nSource: `%s`.\n" % rep)
seen = set()
for b in blocks:
    rows = list(csv.reader(io.StringIO("\n".join(b["rows"]))))
    hdr = rows[0]
    iS, iE, iSm = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    mix, samp, tot, ts = collections.Counter(), collections.Counter(), 0, 0
    for r in rows[1:]:
        if len(r) <= iE or not r[iE].isdigit():
            continue
        toks = r[iS].strip().split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        n, k = int(r[iE]), int(r[iSm]) if r[iSm].isdigit() else 0
        mix[op] += n; samp[op] += k; tot += n; ts += k
    name = list(csv.reader([b["name"]]))[0][1].split("(CUtensorMap")[0].split("(const")[0]
    if (name, tot) in seen:
        continue
    seen.add((name, tot))
    print("## `%s`\n\n%d warp instructions, %d stall samples\n\n| opcode | executed | share | stall samples |\n|---|---:|---:|---:|" % (name, tot, ts))
    for op, n in mix.most_common(N):
        print("| `%s` | %d | %.1f%% | %.1f%% |" % (op, n, 100.0 * n / tot, 100.0 * samp[op] / max(ts, 1)))
    print()
