"""Text Gantt chart of one profiled step (BSPB200_PROFILE_TIMELINE=1 python bench.py ... > line.json):
python tools/timeline_summary.py line.json [ms per character]"""
import json
import sys

NAMES = ["gemm", "potrf_block", "trsm_block", "elim_factor", "elim_gather", "assemble", "solve_elim", "solve_dense", "other", "lump_chol"]
SYM = "gptfGasSoL"
d = json.load(open(sys.argv[1]))
tl = d["kernel_classes"]["timeline"]
res = float(sys.argv[2]) if len(sys.argv) > 2 else 0.25
end = max(r[3] for r in tl)
streams = sorted({r[1] for r in tl})
print(f"records {len(tl)}, span {end:.2f} ms, streams {len(streams)}; legend:", ", ".join(f"{s}={n}" for s, n in zip(SYM, NAMES)))
for si in streams:
    recs = [r for r in tl if r[1] == si]
    busy = sum(r[3] - r[2] for r in recs)
    line = [" "] * (int(end / res) + 1)
    for r in recs:
        for k in range(int(r[2] / res), int(r[3] / res) + 1):
            line[k] = SYM[r[0]]
    print(f"s{si:02d} busy {busy:6.2f} ms n={len(recs):4d} |{''.join(line)}|")
# concurrency profile: number of records running per time slot
slots = [0] * (int(end / res) + 1)
for r in tl:
    for k in range(int(r[2] / res), int(r[3] / res) + 1):
        slots[k] += 1
print("running   " + " " * 22 + "|" + "".join(str(min(9, x)) for x in slots) + "|")
big = sorted(tl, key=lambda r: r[2] - r[3])[:12]
print("longest records:", [(NAMES[r[0]], r[1], round(r[2], 2), round(r[3] - r[2], 2), round(r[4] / 1e9, 2)) for r in big])
