"""Print the tail of an ncu launch list (--metrics gpu__time_duration.sum --csv): kernel, us.
Usage: python profiles/launch_list.py <csv> [n_last]"""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
for r in rows[1:][-n:]:
    print("%-64s %9.1f us" % (r[ki][:64], float(r[vi].replace(",", "")) / 1000))
