"""Per-CUDA-source-line totals (warp instructions executed, stall samples, L1 sectors) of one captured kernel:
python tools/ncu_lines.py file.ncu-rep <launch index> [top]"""
import csv
import subprocess
import sys


def main():
    path, kid = sys.argv[1], int(sys.argv[2])
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(kid),
                          "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    agg = {}
    cur_file, hdr = None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if r[0] == "Function Name" or hdr is None or len(r) != len(hdr):
            continue
        ix = {}
        for i, h in enumerate(hdr):
            ix.setdefault(h, i)
        line = r[0]
        src = r[1]
        key = (cur_file, line)
        a = agg.setdefault(key, {"src": src, "inst": 0.0, "samples": 0.0, "ldsec": 0.0})
        if src.strip():
            a["src"] = src
        try:
            a["inst"] += float(r[ix["Instructions Executed"]] or 0)
            a["samples"] += float(r[ix["# Samples"]] or 0)
        except (ValueError, KeyError):
            pass
    ti = sum(a["inst"] for a in agg.values()) or 1
    ts = sum(a["samples"] for a in agg.values()) or 1
    print("warp instructions %.0f, stall samples %.0f" % (ti, ts))
    for (f, line), a in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[:top]:
        print("%5.1f%% samples %5.1f%% instr  %s:%s  %s" % (100 * a["samples"] / ts, 100 * a["inst"] / ti, f, line, a["src"].strip()[:100]))


if __name__ == "__main__":
    main()
