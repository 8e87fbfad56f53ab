"""Top source lines of one kernel in an ncu report by warp-stall samples (needs -lineinfo and --import-source on).
    python tools/ncu_lines.py report.ncu-rep kernel_regex [top_n] [launch_index]"""
import csv
import io
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 30
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                      f"regex:{pat}"], capture_output=True, text=True).stdout
# the report repeats per launch ("Kernel Name" header lines); keep launch `which`
chunks = out.split('"Kernel Name"')
text = '"Kernel Name"' + chunks[1 + which] if len(chunks) > 1 + which else out
rows = list(csv.reader(io.StringIO(text)))
fname, hdr, data = "", None, []
for r in rows:
    if len(r) >= 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]
    elif len(r) > 4 and r[0] == "Line No":
        hdr = r
    elif hdr and len(r) > 8 and r[0] not in ("", "Line No"):
        try:
            a = hdr.index("Warp Stall Sampling (All Samples)")
            i = hdr.index("Instructions Executed")
            stalls = {h: int(v) for h, v in zip(hdr, r) if h.startswith("stall_") and "Not Issued" not in h and v.isdigit() and int(v) > 0}
            data.append((int(r[a]), int(r[i]), fname, r[0], r[1].strip()[:110], stalls))
        except ValueError:
            pass
tot = sum(d[0] for d in data) or 1
print(f"total samples {tot}, total warp instructions {sum(d[1] for d in data)}")
for d in sorted(data, key=lambda t: -t[0])[:top_n]:
    st = ",".join(f"{k[6:]}:{v}" for k, v in sorted(d[5].items(), key=lambda kv: -kv[1])[:3])
    print(f"{d[0]:6d} {100 * d[0] / tot:5.1f}%  inst {d[1]:8d}  {d[2]}:{d[3]:>4s}  {d[4]}   [{st}]")
