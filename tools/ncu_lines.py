#!/usr/bin/env python
"""Per CUDA source line: warp instructions executed and stall samples of the kernels in an `ncu --set full --import-source on` report.
usage: ncu_lines.py report.ncu-rep [kernel-substring] [top=25]"""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
per_kernel = collections.OrderedDict()
kernel = fname = None
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        kernel = r[1]; continue
    if r[0] == "Line No":
        hdr = r; ii = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); continue
    if hdr is None or r[0] == "" or kernel is None:
        continue
    try:
        inst = float(r[ii]); samp = float(r[isamp] or 0)
    except ValueError:
        continue
    rows_k = per_kernel.setdefault(kernel, collections.OrderedDict())  # a line inlined at several places appears once per place: sum them
    key = (fname, r[0])
    prev = rows_k.get(key, (0.0, 0.0, fname, r[0], r[1].strip()))
    rows_k[key] = (prev[0] + inst, prev[1] + samp, fname, r[0], r[1].strip())
for kernel, data in per_kernel.items():
    data = list(data.values())
    if want not in kernel:
        continue
    tot = sum(d[0] for d in data); tots = sum(d[1] for d in data)
    if tot <= 0:
        continue
    print(f"== {kernel[:90]}: {tot/1e6:.2f} M warp-instr, {int(tots)} samples")
    for d in sorted(data, key=lambda d: -d[0])[:top]:
        print(f"  {d[2][:22]:22s} L{d[3]:>5} {d[0]/tot*100:5.1f}% inst {d[1]/max(tots,1)*100:5.1f}% stall | {d[4][:100]}")
