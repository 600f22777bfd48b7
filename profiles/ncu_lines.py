"""Instructions executed and stall samples per SOURCE LINE of the first profiled launch of an
.ncu-rep (needs -lineinfo and --import-source on):  python profiles/ncu_lines.py rep [top N]"""
import csv
import subprocess
import sys


def main(path, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # several launches / files follow each other: take blocks headed by "Line No"
    agg = {}
    hdr = None
    fname = "?"
    launches = 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            launches += 1
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or launches > 1 and False:
            continue
        try:
            line = int(r[0])
        except ValueError:
            continue
        # the source text may hold quotes and commas that split it into several fields: index the
        # metric columns from the right
        H = len(hdr)
        col = lambda name: r[len(r) - (H - hdr.index(name))]
        key = (fname, line, r[1][:110])
        try:
            inst = int(col("Instructions Executed") or 0)
            samp = int(col("# Samples") or 0)
            thr = int(col("Thread Instructions Executed") or 0)
        except ValueError:
            continue
        a = agg.setdefault(key, [0, 0, 0])
        a[0] += inst
        a[1] += samp
        a[2] += thr
    tot = sum(v[0] for v in agg.values()) or 1
    tots = sum(v[1] for v in agg.values()) or 1
    print(f"total instructions {tot}, samples {tots} (summed over the profiled launches in the report)")
    for (f, line, src), (inst, samp, thr) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * inst / tot:5.1f}% inst {100 * samp / tots:5.1f}% stall  lanes {thr / max(inst, 1):4.1f}  {f}:{line}  {src.strip()}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
