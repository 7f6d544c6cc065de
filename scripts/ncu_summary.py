"""ncu report -> 'metric,unit,value' summary of the first kernel in it (the format of profiles/*/ncu_*_summary.csv)."""
import csv, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
with open(out, "w") as f:
    for h, u, v in zip(hdr, units, vals):
        f.write("%s,%s,%s\n" % (h.replace(",", ";"), u, v.replace(",", "")))
print("wrote", out, len(hdr), "metrics")
