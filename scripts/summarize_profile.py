"""Turn ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

  python scripts/summarize_profile.py launches <launches.csv> <launches_per_eval> <out.md> [title]
  python scripts/summarize_profile.py full <report.ncu-rep> <out.md> [title]
"""
import collections
import csv
import subprocess
import sys


def launches(path, per_eval, out, title):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    rows = [r for r in rows if r["Metric Name"] == "gpu__time_duration.sum"][-per_eval:]
    agg = collections.OrderedDict()
    tot = 0.0
    for r in rows:
        k = r["Kernel Name"].split("(")[0].replace("void ", "").replace("gpb::", "")
        key = (k, r["Grid Size"], r["Block Size"])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"].replace(",", ""))
        tot += float(r["Metric Value"].replace(",", ""))
    byk = collections.defaultdict(lambda: [0, 0.0])
    for (k, g, b), (c, v) in agg.items():
        kk = k.split("<")[0]
        byk[kk][0] += c
        byk[kk][1] += v
    with open(out, "w") as f:
        f.write("# %s\n\n" % title)
        f.write("Source: `ncu --metrics gpu__time_duration.sum --clock-control none` launch list (cold-cache, serialised: compare "
                "SHARES, not absolutes), last %d launches = one NLL+grad evaluation.\n\n" % per_eval)
        f.write("Total kernel time of the evaluation: **%.3f ms** over %d launches.\n\n" % (tot / 1e6, len(rows)))
        f.write("## Share by kernel\n\n| kernel | launches | total ms | share |\n|---|---:|---:|---:|\n")
        for k, (c, v) in sorted(byk.items(), key=lambda x: -x[1][1]):
            f.write("| `%s` | %d | %.3f | %.1f%% |\n" % (k, c, v / 1e6, 100 * v / tot))
        f.write("\n## By kernel instantiation and grid\n\n| kernel | grid | block | launches | total ms | each us |\n|---|---|---|---:|---:|---:|\n")
        for (k, g, b), (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("| `%s` | %s | %s | %d | %.3f | %.1f |\n" % (k, g, b, c, v / 1e6, v / c / 1e3))


WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_bytes.sum",
        "SM_C.TriageCompute.smsp__pipe_tensor_subpipe_dmma_cycles_active.avg",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "smsp__cycles_active.avg"]


def full(path, out, title):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, body = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write("# %s\n\nSource: `ncu --set full --clock-control none --import-source on` (%s).\n\n" % (title, path))
        f.write("| metric | unit | " + " | ".join("launch %d" % i for i in range(len(body))) + " |\n")
        f.write("|---|---|" + "---|" * len(body) + "\n")
        for w in ["Kernel Name", "Grid Size", "Block Size"] + WANT:
            if w in idx:
                f.write("| %s | %s | %s |\n" % (w, units[idx[w]], " | ".join(r[idx[w]].replace("|", "/")[:70] for r in body)))
        src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        srows = list(csv.reader(src.splitlines()))
        secs = [i for i, r in enumerate(srows) if r and r[0] == "Kernel Name"] + [len(srows)]
        for si in range(len(secs) - 1):
            h = srows[secs[si] + 1]
            ix = {x: i for i, x in enumerate(h)}
            b = srows[secs[si] + 2:secs[si + 1]]
            stalls = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
            tot = collections.Counter()
            for r in b:
                for s in stalls:
                    try:
                        tot[s] += int(r[ix[s]])
                    except Exception:
                        pass
            T = sum(tot.values()) or 1
            f.write("\nWarp-stall samples, launch %d (`%s`): " % (si, srows[secs[si]][1][:60]))
            f.write(", ".join("%s %.1f%%" % (k, 100.0 * v / T) for k, v in tot.most_common(7)) + "\n")


if __name__ == "__main__":
    mode = sys.argv[1]
    if mode == "launches":
        launches(sys.argv[2], int(sys.argv[3]), sys.argv[4], sys.argv[5] if len(sys.argv) > 5 else "ncu launch list")
    else:
        full(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "ncu full capture")
