"""Turn an ncu launch list (`--metrics gpu__time_duration.sum --csv`) into a per-kernel table of one bench step.
usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/xxx.md"""
import collections
import csv
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    seq = []
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0].replace("void ", "").replace("BaSpaCho::b200::", "").replace("<unnamed>::", "")
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else v * 1000 if unit == "ms" else v
        seq.append((name, row["Grid Size"], row["Block Size"], v))
    starts = [i for i, s in enumerate(seq) if "elim_factor" in s[0]] or [0]
    step = seq[starts[0]:starts[1]] if len(starts) > 1 else seq
    tot, cnt = collections.Counter(), collections.Counter()
    for n, g, b, v in step:
        tot[n] += v
        cnt[n] += 1
    total = sum(tot.values())
    print(f"launches in one factor()+solve() step: {len(step)}, sum of kernel durations {total / 1000:.3f} ms "
          "(ncu: cold cache, serialised - compare SHARES)\n")
    print("| kernel | launches | total us | share |")
    print("|---|---:|---:|---:|")
    for n, v in tot.most_common():
        print(f"| `{n}` | {cnt[n]} | {v:.1f} | {100 * v / total:.1f}% |")


if __name__ == "__main__":
    main(sys.argv[1])
