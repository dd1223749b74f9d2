"""Per-kernel durations of the last learn step in an ncu launch list (gpu__time_duration.sum)."""
import csv, re, sys
rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
names = [r['Kernel Name'] for r in rows]
key = sys.argv[2] if len(sys.argv) > 2 else 'frames_u8_to_bf16'
idx = [i for i, n in enumerate(names) if key in n]
a, b = idx[-2], idx[-1]
tot = 0
for r in rows[a:b]:
    n = re.sub(r'^void ', '', re.sub(r'\(.*', '', r['Kernel Name']))[:75]; t = float(r['Metric Value']) / 1e3; tot += t
    print(f"{t:8.1f} {r['Grid Size']:>14s} {n}")
print(f"sum {tot:.1f} us over {b - a} launches")
