"""Print the last step's launches (name, us, grid) from an ncu gpu__time_duration CSV."""
import csv, sys
path = sys.argv[1]; per_step = int(sys.argv[2]) if len(sys.argv) > 2 else 26
lines = [l for l in open(path) if not l.startswith('==')]
rows = [(int(x['ID']), x['Kernel Name'].split('(')[0].replace('team::', '').replace('void ', ''), float(x['Metric Value']), x['Grid Size'])
        for x in csv.DictReader(lines) if x.get('Metric Name') == 'gpu__time_duration.sum']
seq = rows[-per_step:]
tot = sum(r[2] for r in seq)
for r in seq:
    print(f"{r[0]:5d} {r[1][:34]:34s} {r[2]/1000:8.2f} us  {r[3]}")
print(f"sum {tot/1000:.1f} us over {len(seq)} launches")
