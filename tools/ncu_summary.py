#!/usr/bin/env python
"""Summaries committed under profiles/ from the raw ncu outputs in gpurun_out/:
    python tools/ncu_summary.py launches <launch-list.csv> <out.txt>      (ncu --metrics gpu__time_duration.sum --csv)
    python tools/ncu_summary.py full <report.ncu-rep> <out.txt>           (ncu --set full)"""
import collections, csv, re, subprocess, sys


def launches(src, out):
    rows = [r for r in csv.reader(open(src, errors="ignore")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    t = collections.defaultdict(list)
    for r in rows:
        if r is hdr or len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[ik]).split("::")[-1][:48]
        t[name].append(float(r[iv].replace(",", "")) / 1e3)   # ns -> us
    tot = sum(sum(v) for v in t.values())
    with open(out, "w") as f:
        f.write("share of the captured window, launches, average duration (cold-cache, serialised under ncu), kernel\n")
        for k, v in sorted(t.items(), key=lambda kv: -sum(kv[1])):
            f.write("%6.2f%% %4d launches %9.1f us avg  %s\n" % (100 * sum(v) / tot, len(v), sum(v) / len(v), k))


def full(rep, out):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[0]
    want = ["gpu__time_duration.sum", "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
    idx = {h: i for i, h in enumerate(hdr)}
    units = rows[1]
    with open(out, "w") as f:
        f.write("ncu --set full --clock-control none, one launch per kernel (values as exported by `ncu --page raw --csv`)\n")
        for r in rows[2:]:
            f.write("\n== %s\n" % r[idx["Kernel Name"]][:100])
            for w in want:
                if w in idx:
                    f.write("  %-80s %s %s\n" % (w, r[idx[w]], units[idx[w]]))
            for h in hdr:
                if "tensor" in h and ".avg.pct" in h and "realtime" not in h and h not in want:
                    f.write("  %-80s %s %s\n" % (h, r[idx[h]], units[idx[h]]))


def source(rep, out, kernels=("tc_field_fwd", "tc_dgrad", "tc_wgrad"), top=25):
    """warp-stall samples per CUDA source line (needs --import-source on and -lineinfo)"""
    with open(out, "w") as f:
        f.write("warp stall samples per source line (ncu --set full --import-source on; top %d lines per kernel)\n" % top)
        for k in kernels:
            txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + k],
                                 capture_output=True, text=True).stdout
            cur, hdr, agg, src = None, None, collections.Counter(), {}
            for r in csv.reader(txt.splitlines()):
                if len(r) == 2 and r[0] == "File Path":
                    cur = r[1].split("/")[-1]
                elif len(r) > 5 and r[0] == "Line No":
                    hdr = r; iS = hdr.index("# Samples")
                elif hdr and len(r) > iS and r[0] != "" and r[2] == "-":
                    try:
                        key = (cur, int(r[0])); agg[key] += int(r[iS]); src[key] = r[1]
                    except ValueError:
                        pass
            tot = sum(agg.values())
            if not tot:
                continue
            f.write("\n== %s (%d samples)\n" % (k, tot))
            for (fn, ln), c in agg.most_common(top):
                f.write("  %5.1f%%  %s:%d  %s\n" % (100.0 * c / tot, fn, ln, src[(fn, ln)].strip()[:110]))


if __name__ == "__main__":
    {"launches": launches, "full": full, "source": source}[sys.argv[1]](sys.argv[2], sys.argv[3])
