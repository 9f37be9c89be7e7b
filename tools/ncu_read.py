"""Reads an .ncu-rep (ncu -i ... --page raw/source --csv) and prints, per captured kernel: duration,
DRAM bytes, throughputs, occupancy, and the SASS instructions holding the most warp-stall samples with
their dominant stall reason. Usage: python tools/ncu_read.py file.ncu-rep [top]"""
import csv
import subprocess
import sys


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    return hdr, rows[2:]


def source(path, k):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-id", "::%s:" % k],
                         capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_lg.sum", "sm__cycles_active.avg", "l1tex__f_wavefronts.sum",
        "l1tex__lsuin_requests.sum"]


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 22
    hdr, rows = raw(path)
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows:
        print("=" * 100)
        for w in WANT:
            if w in ix:
                print("  %-62s %s" % (w, r[ix[w]][:120]))
    for kid in range(len(rows)):
        out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--launch-skip", str(kid), "--launch-count", "1"],
                             capture_output=True, text=True).stdout
        srows = list(csv.reader(out.splitlines()))
        # find header row
        hi = [i for i, r in enumerate(srows) if r and r[0] == "Address"]
        if not hi:
            continue
        h = srows[hi[0]]
        data = [r for r in srows[hi[0] + 1:] if len(r) == len(h)]
        sx = {n: i for i, n in enumerate(h)}
        tot = sum(int(r[sx["# Samples"]]) for r in data) or 1
        stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
        print("-" * 100)
        print(srows[0][1][:110] if srows and len(srows[0]) > 1 else "", "samples", tot, "instructions", len(data))
        agg = {}
        for r in data:
            for n in stalls:
                agg[n] = agg.get(n, 0) + int(r[sx[n]])
        print("  stall totals:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:7]))
        for r in sorted(data, key=lambda r: -int(r[sx["# Samples"]]))[:top]:
            sm = int(r[sx["# Samples"]])
            st = sorted(((int(r[sx[n]]), n[6:]) for n in stalls), reverse=True)[:2]
            print("  %5.1f%% %-64s exec=%-9s %s" % (100.0 * sm / tot, r[sx["Source"]].strip()[:64], r[sx["Instructions Executed"]], st))


if __name__ == "__main__":
    main()
