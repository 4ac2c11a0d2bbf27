"""Print an ncu --csv launch list (gpu__time_duration) as kernel / microseconds / grid."""
import csv, sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
tot = 0.0
for row in csv.DictReader(lines):
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    v = v / 1000 if u == 'ns' else v * 1000 if u == 'ms' else v
    tot += v
    print("%-44s %10.1f us  grid %s" % (row['Kernel Name'][:44], v, row.get('Grid Size')))
print("total %.1f us" % tot)
