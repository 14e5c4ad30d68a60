"""SASS listing of one kernel from an .ncu-rep with executed-instruction counts and stall samples per instruction.
usage: ncu_sass.py report.ncu-rep [min_count] > out.txt"""
import csv
import subprocess
import sys

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stdout=subprocess.PIPE,
                     stderr=subprocess.DEVNULL, text=True).stdout
hdr = None
for r in csv.reader(out.splitlines()):
    if r and r[0] == "Address":
        hdr = r
        continue
    if r and r[0] == "Kernel Name":
        print("==", r[1])
        continue
    if hdr and len(r) >= len(hdr) - 2:
        d = dict(zip(hdr, r))
        print(d["Address"][-5:], d["Source"][:90].ljust(90), d["Instructions Executed"].rjust(10), d["# Samples"].rjust(6))
